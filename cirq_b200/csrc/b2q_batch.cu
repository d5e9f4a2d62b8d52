// Batched Monte-Carlo trajectories: B = 2^b independent n-qubit state vectors
// stored back to back (trajectory t = index bits [n, n+b)), advanced together.
//
// The reference runs a noisy / mid-circuit-measured circuit once per repetition
// (cirq-core/cirq/sim/simulator_base.py:249-264), each time applying, per
// stochastic operation, ONE randomly chosen operator to ONE small state
// (sim/state_vector_simulation_state.py:183-203 mixtures, :205-257 Kraus
// channels, sim/state_vector.py:300-318 measurement collapse).  Here unitary
// gates are ordinary gate passes over the (n+b)-bit array (b2q_apply.cu), and
// the three kernels below do the per-trajectory part of the stochastic steps:
//
//   bsv_apply_select   psi_t <- scale_t * M[choice_t] psi_t   (skipping identity picks)
//   bsv_kraus_weights  w[t][i] = || K_i psi_t ||^2             (read-only)
//   bsv_collapse       psi_t <- scale_t * P(bits == pattern_t) psi_t
//
// All three stream the array once; choices, scales and patterns are drawn on
// the host from the simulator's RandomState and passed as device arrays.
#include "b2q_common.cuh"

#include <algorithm>
#include <vector>

namespace b2q {

constexpr int kSelMaxK = 4;
constexpr int kSelMaxCount = 1 << 16;

struct SelParams {
  int n;               // qubits per trajectory
  int tpos[kSelMaxK];  // ascending target positions (< n)
  int count;           // matrices in the table
  int skip;            // choice value that means "leave the trajectory alone", or -1
};

template <typename real, int K>
__global__ void __launch_bounds__(256)
    bsv_apply_select_kernel(typename Cplx<real>::type* __restrict__ state, uint64_t total_groups,
                            const typename Cplx<real>::type* __restrict__ mats,
                            const int* __restrict__ choice, const double* __restrict__ scale,
                            const SelParams p) {
  using C = typename Cplx<real>::type;
  constexpr int D = 1 << K;
  const int gl2 = p.n - K;  // log2 groups per trajectory
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total_groups; g += stride) {
    const uint64_t t = g >> gl2;
    const int c = choice[t];
    if (c == p.skip) continue;
    const uint64_t r = g & ((1ull << gl2) - 1ull);
    C* base = state + (t << p.n) + insert_zero_bits(r, p.tpos, K);
    C x[D];
#pragma unroll
    for (int j = 0; j < D; ++j) {
      uint64_t off = 0;
#pragma unroll
      for (int b = 0; b < K; ++b)
        if ((j >> b) & 1) off += 1ull << p.tpos[b];
      x[j] = base[off];
    }
    const C* m = mats + (size_t)c * D * D;
    const real s = scale != nullptr ? (real)scale[t] : (real)1;
#pragma unroll
    for (int row = 0; row < D; ++row) {
      C acc = make_c<real>(0, 0);
#pragma unroll
      for (int col = 0; col < D; ++col) {
        const C e = m[row * D + col];
        cmac<real, C>(acc, e.x, e.y, x[col]);
      }
      uint64_t off = 0;
#pragma unroll
      for (int b = 0; b < K; ++b)
        if ((row >> b) & 1) off += 1ull << p.tpos[b];
      base[off] = make_c<real>(acc.x * s, acc.y * s);
    }
  }
}

// grid = (chunks, trajectories): blockIdx.y = t.  out[t * count + i] += partial sums.
template <typename real, int K>
__global__ void __launch_bounds__(256)
    bsv_kraus_weights_kernel(const typename Cplx<real>::type* __restrict__ state,
                             const typename Cplx<real>::type* __restrict__ mats,
                             double* __restrict__ out, uint64_t traj0, const SelParams p) {
  using C = typename Cplx<real>::type;
  constexpr int D = 1 << K;
  extern __shared__ double s_acc[];  // [count]
  const uint64_t t = traj0 + blockIdx.y;
  const uint64_t groups = 1ull << (p.n - K);
  for (int i = threadIdx.x; i < p.count; i += blockDim.x) s_acc[i] = 0.0;
  __syncthreads();
  const C* traj = state + (t << p.n);
  // every thread keeps its own partial sum per operator, a few operators at a time
  for (int i0 = 0; i0 < p.count; i0 += 4) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < groups;
         r += (uint64_t)gridDim.x * blockDim.x) {
      const C* base = traj + insert_zero_bits(r, p.tpos, K);
      C x[D];
#pragma unroll
      for (int j = 0; j < D; ++j) {
        uint64_t off = 0;
#pragma unroll
        for (int b = 0; b < K; ++b)
          if ((j >> b) & 1) off += 1ull << p.tpos[b];
        x[j] = base[off];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (i0 + u >= p.count) break;
        const C* m = mats + (size_t)(i0 + u) * D * D;
        real w = 0;
#pragma unroll
        for (int row = 0; row < D; ++row) {
          C y = make_c<real>(0, 0);
#pragma unroll
          for (int col = 0; col < D; ++col) {
            const C e = m[row * D + col];
            cmac<real, C>(y, e.x, e.y, x[col]);
          }
          w += y.x * y.x + y.y * y.y;
        }
        acc[u] += (double)w;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      double v = acc[u];
      for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
      if ((threadIdx.x & 31) == 0 && i0 + u < p.count) atomicAdd(&s_acc[i0 + u], v);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < p.count; i += blockDim.x)
    atomicAdd(&out[t * p.count + i], s_acc[i]);
}

template <typename real>
__global__ void __launch_bounds__(256)
    bsv_collapse_kernel(typename Cplx<real>::type* __restrict__ state, uint64_t total, int n,
                        uint64_t mask, const uint64_t* __restrict__ pattern,
                        const double* __restrict__ scale) {
  using C = typename Cplx<real>::type;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t local_mask = (1ull << n) - 1ull;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const uint64_t t = i >> n;
    const bool keep = ((i & local_mask) & mask) == pattern[t];
    if (keep) {
      const real s = (real)scale[t];
      C v = state[i];
      v.x *= s;
      v.y *= s;
      state[i] = v;
    } else {
      state[i] = make_c<real>(0, 0);
    }
  }
}

// Several 1-qubit selections in ONE launch: operator j acts on bit tpos[j] and
// trajectory t applies mats[choice[j * B + t]] there (or nothing when the choice
// is `skip`).  One CTA walks a trajectory through its operators in order; a
// noise model's per-moment layer of identical weak channels (most draws are the
// identity) becomes one launch that only touches the trajectories hit.
struct MultiParams {
  int n;
  int m;
  int tpos[32];
  int skip;
};

template <typename real>
__global__ void __launch_bounds__(256)
    bsv_select_multi_kernel(typename Cplx<real>::type* __restrict__ state, uint64_t trajectories,
                            const typename Cplx<real>::type* __restrict__ mats,
                            const int* __restrict__ choice, const __grid_constant__ MultiParams p) {
  using C = typename Cplx<real>::type;
  const uint64_t pairs = 1ull << (p.n - 1);
  for (uint64_t t = blockIdx.x; t < trajectories; t += gridDim.x) {
    C* traj = state + (t << p.n);
    for (int j = 0; j < p.m; ++j) {
      const int c = choice[(uint64_t)j * trajectories + t];  // uniform over the CTA
      if (c == p.skip) continue;
      const C m00 = mats[4 * c], m01 = mats[4 * c + 1], m10 = mats[4 * c + 2],
              m11 = mats[4 * c + 3];
      const int pos = p.tpos[j];
      for (uint64_t g = threadIdx.x; g < pairs; g += blockDim.x) {
        const uint64_t i0 = insert_zero_bit(g, pos);
        const uint64_t i1 = i0 | (1ull << pos);
        const C a = traj[i0], b = traj[i1];
        C r0 = cmul<real, C>(m00.x, m00.y, a);
        cmac<real, C>(r0, m01.x, m01.y, b);
        C r1 = cmul<real, C>(m10.x, m10.y, a);
        cmac<real, C>(r1, m11.x, m11.y, b);
        traj[i0] = r0;
        traj[i1] = r1;
      }
      __syncthreads();  // the next operator reads what this one wrote
    }
  }
}

// matrices (count x 2^k x 2^k complex128, first target = most significant index
// bit) -> device table in sorted-target index order, element type of the state.
template <typename real>
static int upload_matrices(const double* m128, int count, const int* targets, const int* sorted,
                           int k, cudaStream_t stream, typename Cplx<real>::type** out_dev) {
  using C = typename Cplx<real>::type;
  const int d = 1 << k;
  int rank_of[kSelMaxK];  // position of targets[q] in the ascending order
  for (int q = 0; q < k; ++q)
    for (int i = 0; i < k; ++i)
      if (sorted[i] == targets[q]) rank_of[q] = i;
  auto to_orig = [&](int idx_sorted) {
    int o = 0;
    for (int q = 0; q < k; ++q)
      if ((idx_sorted >> rank_of[q]) & 1) o |= 1 << (k - 1 - q);
    return o;
  };
  std::vector<C> host((size_t)count * d * d);
  for (int c = 0; c < count; ++c)
    for (int r = 0; r < d; ++r)
      for (int col = 0; col < d; ++col) {
        const size_t src = ((size_t)c * d * d + (size_t)to_orig(r) * d + to_orig(col)) * 2;
        host[(size_t)c * d * d + r * d + col] = make_c<real>((real)m128[src], (real)m128[src + 1]);
      }
  void* dev = workspace(host.size() * sizeof(C));
  if (dev == nullptr) return B2Q_ERR_CUDA;
  B2Q_CUDA_CHECK(cudaMemcpyAsync(dev, host.data(), host.size() * sizeof(C), cudaMemcpyHostToDevice,
                                 stream));
  // (a copy from pageable memory returns once the source has been staged, so the
  // host vector may die at return)
  *out_dev = reinterpret_cast<C*>(dev);
  return B2Q_OK;
}

static int check_batch_args(int dtype, int n, int b, const int* targets, int k, int* sorted) {
  B2Q_REQUIRE(dtype == B2Q_C64 || dtype == B2Q_C128, "bad dtype %d", dtype);
  B2Q_REQUIRE(n >= 1 && b >= 0 && n + b <= 40, "bad shape: %d qubits, %d batch bits", n, b);
  B2Q_REQUIRE(k >= 1 && k <= kSelMaxK && k <= n, "operators of 1..%d qubits, got %d", kSelMaxK, k);
  for (int i = 0; i < k; ++i) sorted[i] = targets[i];
  std::sort(sorted, sorted + k);
  for (int i = 0; i < k; ++i) {
    B2Q_REQUIRE(sorted[i] >= 0 && sorted[i] < n, "target bit %d out of range for %d qubits",
                sorted[i], n);
    B2Q_REQUIRE(i == 0 || sorted[i] != sorted[i - 1], "duplicate target bit %d", sorted[i]);
  }
  return B2Q_OK;
}

template <typename real>
static int apply_select_t(void* state, int n, int b, const double* m128, int count,
                          const int* targets, const int* sorted, int k, const int* choice,
                          const double* scale, int skip, cudaStream_t stream) {
  using C = typename Cplx<real>::type;
  C* mats = nullptr;
  const int rc = upload_matrices<real>(m128, count, targets, sorted, k, stream, &mats);
  if (rc != B2Q_OK) return rc;
  SelParams p;
  p.n = n;
  for (int i = 0; i < kSelMaxK; ++i) p.tpos[i] = i < k ? sorted[i] : 0;
  p.count = count;
  p.skip = skip;
  const uint64_t total_groups = 1ull << (n + b - k);
  const unsigned blocks =
      (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((total_groups + 255) / 256, 148ull * 32));
  C* s = reinterpret_cast<C*>(state);
  switch (k) {
    case 1: bsv_apply_select_kernel<real, 1><<<blocks, 256, 0, stream>>>(s, total_groups, mats, choice, scale, p); break;
    case 2: bsv_apply_select_kernel<real, 2><<<blocks, 256, 0, stream>>>(s, total_groups, mats, choice, scale, p); break;
    case 3: bsv_apply_select_kernel<real, 3><<<blocks, 256, 0, stream>>>(s, total_groups, mats, choice, scale, p); break;
    default: bsv_apply_select_kernel<real, 4><<<blocks, 256, 0, stream>>>(s, total_groups, mats, choice, scale, p); break;
  }
  B2Q_LAUNCH_CHECK("bsv_apply_select_kernel");
  return B2Q_OK;
}

template <typename real>
static int kraus_weights_t(const void* state, int n, int b, const double* m128, int count,
                           const int* targets, const int* sorted, int k, double* out,
                           cudaStream_t stream) {
  using C = typename Cplx<real>::type;
  C* mats = nullptr;
  const int rc = upload_matrices<real>(m128, count, targets, sorted, k, stream, &mats);
  if (rc != B2Q_OK) return rc;
  SelParams p;
  p.n = n;
  for (int i = 0; i < kSelMaxK; ++i) p.tpos[i] = i < k ? sorted[i] : 0;
  p.count = count;
  p.skip = -1;
  const uint64_t trajectories = 1ull << b;
  B2Q_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(double) * trajectories * count, stream));
  const uint64_t groups = 1ull << (n - k);
  const unsigned chunks = (unsigned)std::max<uint64_t>(
      1, std::min<uint64_t>((groups + 255) / 256, std::max<uint64_t>(1, (148ull * 16) >> std::min(b, 11))));
  const C* s = reinterpret_cast<const C*>(state);
  for (uint64_t t0 = 0; t0 < trajectories; t0 += 32768) {
    const unsigned ny = (unsigned)std::min<uint64_t>(32768, trajectories - t0);
    const dim3 grid(chunks, ny);
    const size_t smem = sizeof(double) * count;
    switch (k) {
      case 1: bsv_kraus_weights_kernel<real, 1><<<grid, 256, smem, stream>>>(s, mats, out, t0, p); break;
      case 2: bsv_kraus_weights_kernel<real, 2><<<grid, 256, smem, stream>>>(s, mats, out, t0, p); break;
      case 3: bsv_kraus_weights_kernel<real, 3><<<grid, 256, smem, stream>>>(s, mats, out, t0, p); break;
      default: bsv_kraus_weights_kernel<real, 4><<<grid, 256, smem, stream>>>(s, mats, out, t0, p); break;
    }
    B2Q_LAUNCH_CHECK("bsv_kraus_weights_kernel");
  }
  return B2Q_OK;
}

}  // namespace b2q

using namespace b2q;

extern "C" int b2q_bsv_apply_select(void* state, int dtype, int n_qubits, int batch_bits,
                                    const double* matrices_c128, int count, const int* targets,
                                    int k, const int* choice_dev, const double* scale_dev,
                                    int skip_index, void* stream) {
  B2Q_REQUIRE(state != nullptr && matrices_c128 != nullptr && targets != nullptr &&
                  choice_dev != nullptr,
              "null argument");
  B2Q_REQUIRE(count >= 1 && count <= kSelMaxCount, "1..%d operators, got %d", kSelMaxCount, count);
  int sorted[kSelMaxK];
  const int rc = check_batch_args(dtype, n_qubits, batch_bits, targets, k, sorted);
  if (rc != B2Q_OK) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return dtype == B2Q_C64
             ? apply_select_t<float>(state, n_qubits, batch_bits, matrices_c128, count, targets,
                                     sorted, k, choice_dev, scale_dev, skip_index, s)
             : apply_select_t<double>(state, n_qubits, batch_bits, matrices_c128, count, targets,
                                      sorted, k, choice_dev, scale_dev, skip_index, s);
}

extern "C" int b2q_bsv_apply_select_multi(void* state, int dtype, int n_qubits, int batch_bits,
                                          const double* matrices_c128, int count,
                                          const int* targets, int m, const int* choice_dev,
                                          int skip_index, void* stream) {
  B2Q_REQUIRE(state != nullptr && matrices_c128 != nullptr && targets != nullptr &&
                  choice_dev != nullptr,
              "null argument");
  B2Q_REQUIRE(dtype == B2Q_C64 || dtype == B2Q_C128, "bad dtype %d", dtype);
  B2Q_REQUIRE(n_qubits >= 1 && batch_bits >= 0 && n_qubits + batch_bits <= 40, "bad shape");
  B2Q_REQUIRE(count >= 1 && count <= kSelMaxCount, "1..%d operators, got %d", kSelMaxCount, count);
  B2Q_REQUIRE(m >= 1 && m <= 32, "1..32 targets, got %d", m);
  MultiParams p;
  p.n = n_qubits;
  p.m = m;
  p.skip = skip_index;
  for (int j = 0; j < 32; ++j) p.tpos[j] = 0;
  for (int j = 0; j < m; ++j) {
    B2Q_REQUIRE(targets[j] >= 0 && targets[j] < n_qubits, "target bit %d out of range", targets[j]);
    p.tpos[j] = targets[j];
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const uint64_t trajectories = 1ull << batch_bits;
  const unsigned blocks = (unsigned)std::min<uint64_t>(trajectories, 148ull * 16);
  const int one = 0;
  if (dtype == B2Q_C64) {
    float2* mats = nullptr;
    const int rc = upload_matrices<float>(matrices_c128, count, &one, &one, 1, s, &mats);
    if (rc != B2Q_OK) return rc;
    bsv_select_multi_kernel<float><<<blocks, 256, 0, s>>>(reinterpret_cast<float2*>(state),
                                                          trajectories, mats, choice_dev, p);
  } else {
    double2* mats = nullptr;
    const int rc = upload_matrices<double>(matrices_c128, count, &one, &one, 1, s, &mats);
    if (rc != B2Q_OK) return rc;
    bsv_select_multi_kernel<double><<<blocks, 256, 0, s>>>(reinterpret_cast<double2*>(state),
                                                           trajectories, mats, choice_dev, p);
  }
  B2Q_LAUNCH_CHECK("bsv_select_multi_kernel");
  return B2Q_OK;
}

extern "C" int b2q_bsv_kraus_weights(const void* state, int dtype, int n_qubits, int batch_bits,
                                     const double* matrices_c128, int count, const int* targets,
                                     int k, double* weights_dev, void* stream) {
  B2Q_REQUIRE(state != nullptr && matrices_c128 != nullptr && targets != nullptr &&
                  weights_dev != nullptr,
              "null argument");
  B2Q_REQUIRE(count >= 1 && count <= 256, "1..256 Kraus operators, got %d", count);
  int sorted[kSelMaxK];
  const int rc = check_batch_args(dtype, n_qubits, batch_bits, targets, k, sorted);
  if (rc != B2Q_OK) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return dtype == B2Q_C64
             ? kraus_weights_t<float>(state, n_qubits, batch_bits, matrices_c128, count, targets,
                                      sorted, k, weights_dev, s)
             : kraus_weights_t<double>(state, n_qubits, batch_bits, matrices_c128, count, targets,
                                       sorted, k, weights_dev, s);
}

extern "C" int b2q_bsv_collapse(void* state, int dtype, int n_qubits, int batch_bits, uint64_t mask,
                                const uint64_t* pattern_dev, const double* scale_dev, void* stream) {
  B2Q_REQUIRE(state != nullptr && pattern_dev != nullptr && scale_dev != nullptr, "null argument");
  B2Q_REQUIRE(dtype == B2Q_C64 || dtype == B2Q_C128, "bad dtype %d", dtype);
  B2Q_REQUIRE(n_qubits >= 1 && batch_bits >= 0 && n_qubits + batch_bits <= 40, "bad shape");
  B2Q_REQUIRE((mask >> n_qubits) == 0, "mask has bits outside the %d qubits", n_qubits);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const uint64_t total = 1ull << (n_qubits + batch_bits);
  const unsigned blocks =
      (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((total + 255) / 256, 148ull * 32));
  if (dtype == B2Q_C64)
    bsv_collapse_kernel<float><<<blocks, 256, 0, s>>>(reinterpret_cast<float2*>(state), total,
                                                      n_qubits, mask, pattern_dev, scale_dev);
  else
    bsv_collapse_kernel<double><<<blocks, 256, 0, s>>>(reinterpret_cast<double2*>(state), total,
                                                       n_qubits, mask, pattern_dev, scale_dev);
  B2Q_LAUNCH_CHECK("bsv_collapse_kernel");
  return B2Q_OK;
}
