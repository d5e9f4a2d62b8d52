// State-layout operations behind split_untangled_states and ancilla qubits:
// Kronecker product, axis (bit) permutation, arg-max |amplitude|, and the
// separability check of factor().
//
// Replaces (reference, cirq-core/cirq/): linalg/transformations.py:603-613
// (state_vector_kronecker_product = np.outer), :743-754
// (transpose_state_vector_to_axis_order = np.moveaxis) and the numeric steps
// of factor_state_vector (:647-691), as used by
// sim/state_vector_simulation_state.py:105-159 (kron / factor / reindex) and
// sim/simulation_product_state.py:68-139.
#include "b2q_common.cuh"

#include <algorithm>
#include <vector>

namespace b2q {

// out[(i << nb) | j] = a[i] * b[j]
template <typename real>
__global__ void __launch_bounds__(256)
    sv_kron_kernel(const typename Cplx<real>::type* __restrict__ a,
                   const typename Cplx<real>::type* __restrict__ b, int nb,
                   typename Cplx<real>::type* __restrict__ out, uint64_t total) {
  using C = typename Cplx<real>::type;
  const uint64_t mask = (1ull << nb) - 1ull;
  for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total;
       o += (uint64_t)gridDim.x * blockDim.x) {
    const C x = a[o >> nb];
    const C y = b[o & mask];
    out[o] = make_c<real>(x.x * y.x - x.y * y.y, x.x * y.y + x.y * y.x);
  }
}

struct PermuteParams {
  int n;
  int src_bit[40];  // output index bit k comes from input index bit src_bit[k]
};

// out[o] = in[i], bit k of o == bit src_bit[k] of i
template <typename real>
__global__ void __launch_bounds__(256)
    sv_permute_kernel(const typename Cplx<real>::type* __restrict__ in,
                      typename Cplx<real>::type* __restrict__ out,
                      const __grid_constant__ PermuteParams p) {
  const uint64_t total = 1ull << p.n;
  for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total;
       o += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t i = 0;
    for (int k = 0; k < p.n; ++k) i |= ((o >> k) & 1ull) << p.src_bit[k];
    out[o] = in[i];
  }
}

// (max |psi|^2, first index attaining it) per CTA, then over CTAs.
template <typename real>
__global__ void __launch_bounds__(256)
    sv_argmax_partial_kernel(const typename Cplx<real>::type* __restrict__ state, uint64_t total,
                             double* __restrict__ best_val, uint64_t* __restrict__ best_idx) {
  double bv = -1.0;
  uint64_t bi = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const auto a = state[i];
    const double v = (double)a.x * (double)a.x + (double)a.y * (double)a.y;
    if (v > bv) {
      bv = v;
      bi = i;
    }
  }
  __shared__ double sv[256];
  __shared__ uint64_t si[256];
  sv[threadIdx.x] = bv;
  si[threadIdx.x] = bi;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      const double ov = sv[threadIdx.x + s];
      const uint64_t oi = si[threadIdx.x + s];
      if (ov > sv[threadIdx.x] || (ov == sv[threadIdx.x] && oi < si[threadIdx.x])) {
        sv[threadIdx.x] = ov;
        si[threadIdx.x] = oi;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    best_val[blockIdx.x] = sv[0];
    best_idx[blockIdx.x] = si[0];
  }
}

// flag[0] = 1 if some |a[i]*b[j] - t[(i<<nb)|j]| > atol + rtol*|t|  (np.allclose)
template <typename real>
__global__ void __launch_bounds__(256)
    sv_kron_mismatch_kernel(const typename Cplx<real>::type* __restrict__ a,
                            const typename Cplx<real>::type* __restrict__ b, int nb,
                            const typename Cplx<real>::type* __restrict__ t, uint64_t total,
                            double atol, double rtol, int* __restrict__ flag) {
  const uint64_t mask = (1ull << nb) - 1ull;
  bool bad = false;
  for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total;
       o += (uint64_t)gridDim.x * blockDim.x) {
    const auto x = a[o >> nb];
    const auto y = b[o & mask];
    const auto z = t[o];
    const double dr = (double)(x.x * y.x - x.y * y.y) - (double)z.x;
    const double di = (double)(x.x * y.y + x.y * y.x) - (double)z.y;
    const double lim = atol + rtol * sqrt((double)z.x * z.x + (double)z.y * z.y);
    if (sqrt(dr * dr + di * di) > lim) bad = true;
  }
  if (bad) atomicExch(flag, 1);
}

struct PartialTraceParams {
  int n;          // qubits of rho
  int k;          // kept qubits
  int keep[20];   // kept column-bit positions, keep[0] = MSB of the output index
  int traced[20]; // traced column-bit positions (any order)
};

// out[(R << k) | C] = sum_x rho[row(R, x), col(C, x)]; one warp per output element.
template <typename real>
__global__ void __launch_bounds__(256)
    dm_partial_trace_kernel(const typename Cplx<real>::type* __restrict__ rho,
                            typename Cplx<real>::type* __restrict__ out,
                            const __grid_constant__ PartialTraceParams p) {
  const int lane = threadIdx.x & 31;
  const uint64_t o = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const uint64_t total_out = 1ull << (2 * p.k);
  if (o >= total_out) return;
  const uint64_t R = o >> p.k, C = o & ((1ull << p.k) - 1ull);
  uint64_t rbase = 0, cbase = 0;
  for (int q = 0; q < p.k; ++q) {
    rbase |= ((R >> (p.k - 1 - q)) & 1ull) << p.keep[q];
    cbase |= ((C >> (p.k - 1 - q)) & 1ull) << p.keep[q];
  }
  const int nt = p.n - p.k;
  double ar = 0.0, ai = 0.0;
  for (uint64_t x = lane; x < (1ull << nt); x += 32) {
    uint64_t dep = 0;
    for (int t = 0; t < nt; ++t) dep |= ((x >> t) & 1ull) << p.traced[t];
    const auto v = rho[((rbase | dep) << p.n) | (cbase | dep)];
    ar += (double)v.x;
    ai += (double)v.y;
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    ar += __shfl_xor_sync(0xffffffffu, ar, s);
    ai += __shfl_xor_sync(0xffffffffu, ai, s);
  }
  if (lane == 0) out[o] = make_c<real>((real)ar, (real)ai);
}

// flag[0] = 1 if some |a[i] - b[i]| > atol + rtol*|b[i]|  (np.allclose(a, b))
template <typename real>
__global__ void __launch_bounds__(256)
    sv_mismatch_kernel(const typename Cplx<real>::type* __restrict__ a,
                       const typename Cplx<real>::type* __restrict__ b, uint64_t total,
                       double atol, double rtol, int* __restrict__ flag) {
  bool bad = false;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const auto x = a[i];
    const auto y = b[i];
    const double dr = (double)x.x - (double)y.x, di = (double)x.y - (double)y.y;
    if (sqrt(dr * dr + di * di) > atol + rtol * sqrt((double)y.x * y.x + (double)y.y * y.y))
      bad = true;
  }
  if (bad) atomicExch(flag, 1);
}

inline unsigned layout_grid(uint64_t total) {
  return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((total + 255) / 256, 148ull * 64));
}

}  // namespace b2q

using namespace b2q;

extern "C" int b2q_sv_kron(const void* a, int na, const void* b, int nb, int dtype, void* out,
                           void* stream) {
  B2Q_REQUIRE(a != nullptr && b != nullptr && out != nullptr, "null argument");
  B2Q_REQUIRE(dtype == B2Q_C64 || dtype == B2Q_C128, "bad dtype %d", dtype);
  B2Q_REQUIRE(na >= 0 && nb >= 0 && na + nb <= 40, "qubit counts out of range");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const uint64_t total = 1ull << (na + nb);
  if (dtype == B2Q_C64)
    sv_kron_kernel<float><<<layout_grid(total), 256, 0, s>>>(
        reinterpret_cast<const float2*>(a), reinterpret_cast<const float2*>(b), nb,
        reinterpret_cast<float2*>(out), total);
  else
    sv_kron_kernel<double><<<layout_grid(total), 256, 0, s>>>(
        reinterpret_cast<const double2*>(a), reinterpret_cast<const double2*>(b), nb,
        reinterpret_cast<double2*>(out), total);
  B2Q_LAUNCH_CHECK("sv_kron_kernel");
  return B2Q_OK;
}

extern "C" int b2q_sv_permute_bits(const void* in, void* out, int dtype, int n_qubits,
                                          const int* src_bit, void* stream) {
  B2Q_REQUIRE(in != nullptr && out != nullptr && in != out, "bad buffers (must not alias)");
  B2Q_REQUIRE(dtype == B2Q_C64 || dtype == B2Q_C128, "bad dtype %d", dtype);
  B2Q_REQUIRE(n_qubits >= 0 && n_qubits <= 40, "n_qubits out of range");
  PermuteParams p;
  p.n = n_qubits;
  uint64_t seen = 0;
  for (int k = 0; k < n_qubits; ++k) {
    B2Q_REQUIRE(src_bit[k] >= 0 && src_bit[k] < n_qubits, "source bit out of range");
    B2Q_REQUIRE(!((seen >> src_bit[k]) & 1ull), "src_bit is not a permutation");
    seen |= 1ull << src_bit[k];
    p.src_bit[k] = src_bit[k];
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const uint64_t total = 1ull << n_qubits;
  if (dtype == B2Q_C64)
    sv_permute_kernel<float><<<layout_grid(total), 256, 0, s>>>(
        reinterpret_cast<const float2*>(in), reinterpret_cast<float2*>(out), p);
  else
    sv_permute_kernel<double><<<layout_grid(total), 256, 0, s>>>(
        reinterpret_cast<const double2*>(in), reinterpret_cast<double2*>(out), p);
  B2Q_LAUNCH_CHECK("sv_permute_kernel");
  return B2Q_OK;
}

extern "C" int b2q_sv_argmax_abs(const void* state, int dtype, int n_qubits, uint64_t* index_out,
                                 void* stream) {
  B2Q_REQUIRE(state != nullptr && index_out != nullptr, "null argument");
  B2Q_REQUIRE(dtype == B2Q_C64 || dtype == B2Q_C128, "bad dtype %d", dtype);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const uint64_t total = 1ull << n_qubits;
  const unsigned blocks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((total + 255) / 256, 1024));
  double* vals = reinterpret_cast<double*>(workspace(blocks * (sizeof(double) + sizeof(uint64_t))));
  if (vals == nullptr) return B2Q_ERR_CUDA;
  uint64_t* idx = reinterpret_cast<uint64_t*>(vals + blocks);
  if (dtype == B2Q_C64)
    sv_argmax_partial_kernel<float><<<blocks, 256, 0, s>>>(reinterpret_cast<const float2*>(state),
                                                           total, vals, idx);
  else
    sv_argmax_partial_kernel<double><<<blocks, 256, 0, s>>>(
        reinterpret_cast<const double2*>(state), total, vals, idx);
  B2Q_LAUNCH_CHECK("sv_argmax_partial_kernel");
  std::vector<double> hv(blocks);
  std::vector<uint64_t> hi(blocks);
  B2Q_CUDA_CHECK(cudaMemcpyAsync(hv.data(), vals, sizeof(double) * blocks, cudaMemcpyDeviceToHost, s));
  B2Q_CUDA_CHECK(cudaMemcpyAsync(hi.data(), idx, sizeof(uint64_t) * blocks, cudaMemcpyDeviceToHost, s));
  B2Q_CUDA_CHECK(cudaStreamSynchronize(s));
  double bv = -1.0;
  uint64_t bi = 0;
  for (unsigned b = 0; b < blocks; ++b)
    if (hv[b] > bv || (hv[b] == bv && hi[b] < bi)) {
      bv = hv[b];
      bi = hi[b];
    }
  *index_out = bi;
  return B2Q_OK;
}

extern "C" int b2q_sv_kron_allclose(const void* a, int na, const void* b, int nb, const void* t,
                                    int dtype, double atol, double rtol, int* ok_out,
                                    void* stream) {
  B2Q_REQUIRE(a && b && t && ok_out, "null argument");
  B2Q_REQUIRE(dtype == B2Q_C64 || dtype == B2Q_C128, "bad dtype %d", dtype);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const uint64_t total = 1ull << (na + nb);
  int* flag = reinterpret_cast<int*>(workspace(sizeof(int)));
  if (flag == nullptr) return B2Q_ERR_CUDA;
  B2Q_CUDA_CHECK(cudaMemsetAsync(flag, 0, sizeof(int), s));
  if (dtype == B2Q_C64)
    sv_kron_mismatch_kernel<float><<<layout_grid(total), 256, 0, s>>>(
        reinterpret_cast<const float2*>(a), reinterpret_cast<const float2*>(b), nb,
        reinterpret_cast<const float2*>(t), total, atol, rtol, flag);
  else
    sv_kron_mismatch_kernel<double><<<layout_grid(total), 256, 0, s>>>(
        reinterpret_cast<const double2*>(a), reinterpret_cast<const double2*>(b), nb,
        reinterpret_cast<const double2*>(t), total, atol, rtol, flag);
  B2Q_LAUNCH_CHECK("sv_kron_mismatch_kernel");
  int h = 0;
  B2Q_CUDA_CHECK(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, s));
  B2Q_CUDA_CHECK(cudaStreamSynchronize(s));
  *ok_out = h ? 0 : 1;
  return B2Q_OK;
}

extern "C" int b2q_dm_partial_trace(const void* rho, int dtype, int n_qubits, const int* keep_bits,
                                    int k, void* out, void* stream) {
  B2Q_REQUIRE(rho != nullptr && out != nullptr && (k == 0 || keep_bits != nullptr), "null argument");
  B2Q_REQUIRE(dtype == B2Q_C64 || dtype == B2Q_C128, "bad dtype %d", dtype);
  B2Q_REQUIRE(n_qubits >= 0 && n_qubits <= 20 && k >= 0 && k <= n_qubits, "sizes out of range");
  PartialTraceParams p;
  p.n = n_qubits;
  p.k = k;
  uint64_t seen = 0;
  for (int q = 0; q < k; ++q) {
    B2Q_REQUIRE(keep_bits[q] >= 0 && keep_bits[q] < n_qubits, "bit out of range");
    B2Q_REQUIRE(!((seen >> keep_bits[q]) & 1ull), "duplicate bit");
    seen |= 1ull << keep_bits[q];
    p.keep[q] = keep_bits[q];
  }
  int t = 0;
  for (int b = 0; b < n_qubits; ++b)
    if (!((seen >> b) & 1ull)) p.traced[t++] = b;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const uint64_t total_out = 1ull << (2 * k);
  const uint64_t blocks = (total_out + 7) / 8;
  B2Q_REQUIRE(blocks <= 0x7fffffffull, "grid too large");
  if (dtype == B2Q_C64)
    dm_partial_trace_kernel<float><<<(unsigned)blocks, 256, 0, s>>>(
        reinterpret_cast<const float2*>(rho), reinterpret_cast<float2*>(out), p);
  else
    dm_partial_trace_kernel<double><<<(unsigned)blocks, 256, 0, s>>>(
        reinterpret_cast<const double2*>(rho), reinterpret_cast<double2*>(out), p);
  B2Q_LAUNCH_CHECK("dm_partial_trace_kernel");
  return B2Q_OK;
}

extern "C" int b2q_sv_allclose(const void* a, const void* b, int dtype, int n_qubits, double atol,
                               double rtol, int* ok_out, void* stream) {
  B2Q_REQUIRE(a && b && ok_out, "null argument");
  B2Q_REQUIRE(dtype == B2Q_C64 || dtype == B2Q_C128, "bad dtype %d", dtype);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const uint64_t total = 1ull << n_qubits;
  int* flag = reinterpret_cast<int*>(workspace(sizeof(int)));
  if (flag == nullptr) return B2Q_ERR_CUDA;
  B2Q_CUDA_CHECK(cudaMemsetAsync(flag, 0, sizeof(int), s));
  if (dtype == B2Q_C64)
    sv_mismatch_kernel<float><<<layout_grid(total), 256, 0, s>>>(
        reinterpret_cast<const float2*>(a), reinterpret_cast<const float2*>(b), total, atol, rtol,
        flag);
  else
    sv_mismatch_kernel<double><<<layout_grid(total), 256, 0, s>>>(
        reinterpret_cast<const double2*>(a), reinterpret_cast<const double2*>(b), total, atol,
        rtol, flag);
  B2Q_LAUNCH_CHECK("sv_mismatch_kernel");
  int h = 0;
  B2Q_CUDA_CHECK(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, s));
  B2Q_CUDA_CHECK(cudaStreamSynchronize(s));
  *ok_out = h ? 0 : 1;
  return B2Q_OK;
}
