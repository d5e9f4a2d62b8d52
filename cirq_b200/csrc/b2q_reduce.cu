// Read-out side of the hot path: norms, marginal probabilities, inverse-CDF
// sampling, collapse, amplitude gather, Pauli expectation values, and the
// density-matrix diagonal / collapse.
//
// Replaces (reference, cirq-core/cirq/): sim/state_vector.py:170-322
// (sample_state_vector / measure_state_vector), sim/simulation_utils.py:24-65
// (state_probabilities_by_indices), sim/density_matrix_utils.py:31-192,
// ops/pauli_string.py:625-655 and the fancy-index gather of
// sim/state_vector_simulator.py:95-98.
//
// Sampling is a three-level inverse CDF: one streaming
// read of the state produces float64 sums of 128-amplitude chunks, chunks are
// grouped by 128 into blocks, block sums are prefix-scanned, and each sample is
// resolved by one warp: binary search over blocks, warp scan over the block's
// chunk sums, warp scan over the chunk's amplitudes (re-read, ~1.5 KB/sample).
#include "b2q_common.cuh"

#include <algorithm>
#include <cstdlib>
#include <vector>

namespace b2q {

constexpr int kChunkLog2 = 7;  // amplitudes per chunk (<= 128 = 4 per lane)
constexpr int kCpbLog2 = 7;    // chunks per block

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double warp_inclusive_scan(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

template <typename real>
__device__ __forceinline__ double abs2(const typename Cplx<real>::type& a) {
  // |a|^2 in the state's real type, as (psi * psi.conj()).real of the
  // reference (sim/state_vector.py:220), then widened for accumulation.
  return (double)(a.x * a.x + a.y * a.y);
}

// ---- block-wise norm --------------------------------------------------------

// 16-byte streaming access helpers: V16<real> is the 16-byte vector type, holding
// 2 complex64 or 1 complex128 amplitudes.
template <typename real>
struct V16;
template <>
struct V16<float> {
  using type = float4;
  static constexpr int kElems = 2;
  static __device__ __forceinline__ double abs2sum(const float4& v) {
    return (double)(v.x * v.x + v.y * v.y) + (double)(v.z * v.z + v.w * v.w);
  }
};
template <>
struct V16<double> {
  using type = double2;
  static constexpr int kElems = 1;
  static __device__ __forceinline__ double abs2sum(const double2& v) {
    return v.x * v.x + v.y * v.y;
  }
};

constexpr int kStreamUnroll = 4;

// partial[b] = sum of |psi|^2 over a grid-strided subset; deterministic order.
// Each thread keeps kStreamUnroll independent 16-byte loads in flight.
template <typename real>
__global__ void __launch_bounds__(256)
    sv_norm_partial_kernel(const typename Cplx<real>::type* __restrict__ state, uint64_t total,
                           double* __restrict__ partial) {
  using V = typename V16<real>::type;
  const uint64_t nvec = total / V16<real>::kElems;  // total >= 2 for c64 is ensured by the host
  const V* __restrict__ vp = reinterpret_cast<const V*>(state);
  double acc[kStreamUnroll];
#pragma unroll
  for (int u = 0; u < kStreamUnroll; ++u) acc[u] = 0.0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (kStreamUnroll - 1) * stride < nvec; i += kStreamUnroll * stride) {
    V v[kStreamUnroll];
#pragma unroll
    for (int u = 0; u < kStreamUnroll; ++u) v[u] = vp[i + u * stride];
#pragma unroll
    for (int u = 0; u < kStreamUnroll; ++u) acc[u] += V16<real>::abs2sum(v[u]);
  }
  for (; i < nvec; i += stride) acc[0] += V16<real>::abs2sum(vp[i]);
  double a = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  __shared__ double sm[8];
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < 8; ++w) t += sm[w];
    partial[blockIdx.x] = t;
  }
}

// Sums `count` doubles (<= a few thousand) with one CTA into out[0..ncomp).
// in is laid out [count][ncomp].
__global__ void __launch_bounds__(256)
    final_sum_kernel(const double* __restrict__ in, uint64_t count, int ncomp,
                     double* __restrict__ out) {
  __shared__ double sm[8];
  for (int comp = 0; comp < ncomp; ++comp) {
    double acc = 0.0;
    for (uint64_t i = threadIdx.x; i < count; i += blockDim.x) acc += in[i * ncomp + comp];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0;
      for (int w = 0; w < 8; ++w) t += sm[w];
      out[comp] = t;
    }
    __syncthreads();
  }
}

// ---- init / scale / collapse ------------------------------------------------

template <typename real>
__global__ void set_one_kernel(typename Cplx<real>::type* state, uint64_t index) {
  state[index] = make_c<real>(1, 0);
}

template <typename real>
__device__ __forceinline__ typename V16<real>::type scale_v16(const typename V16<real>::type& v,
                                                              real re, real im);
template <>
__device__ __forceinline__ float4 scale_v16<float>(const float4& v, float re, float im) {
  return make_float4(re * v.x - im * v.y, re * v.y + im * v.x, re * v.z - im * v.w,
                     re * v.w + im * v.z);
}
template <>
__device__ __forceinline__ double2 scale_v16<double>(const double2& v, double re, double im) {
  return make_double2(re * v.x - im * v.y, re * v.y + im * v.x);
}

// state <- (re + i im) * state on the amplitudes with (index & mask) == want,
// zero elsewhere (mask == 0: plain scaling).  16-byte vectors, 4 in flight.
template <typename real>
__global__ void __launch_bounds__(256)
    sv_scale_mask_kernel(typename Cplx<real>::type* __restrict__ state, uint64_t total,
                         uint64_t mask, uint64_t want, real re, real im) {
  using V = typename V16<real>::type;
  constexpr int E = V16<real>::kElems;
  const uint64_t nvec = total / E;
  V* __restrict__ vp = reinterpret_cast<V*>(state);
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  auto apply = [&](uint64_t vi, V v) {
    V r = scale_v16<real>(v, re, im);
    if (mask) {
      const uint64_t e0 = vi * E;
      if constexpr (E == 2) {
        if ((e0 & mask) != want) r.x = r.y = 0;
        if (((e0 | 1ull) & mask) != want) r.z = r.w = 0;
      } else {
        if ((e0 & mask) != want) r.x = r.y = 0;
      }
    }
    return r;
  };
  for (; i + (kStreamUnroll - 1) * stride < nvec; i += kStreamUnroll * stride) {
    V v[kStreamUnroll];
#pragma unroll
    for (int u = 0; u < kStreamUnroll; ++u) v[u] = vp[i + u * stride];
#pragma unroll
    for (int u = 0; u < kStreamUnroll; ++u) vp[i + u * stride] = apply(i + u * stride, v[u]);
  }
  for (; i < nvec; i += stride) vp[i] = apply(i, vp[i]);
}

// Scalar fallback for the 1-amplitude complex64 state (n = 0).
template <typename real>
__global__ void sv_scale_tiny_kernel(typename Cplx<real>::type* state, uint64_t total,
                                     uint64_t mask, uint64_t want, real re, real im) {
  for (uint64_t i = threadIdx.x; i < total; i += blockDim.x) {
    if ((i & mask) == want) {
      state[i] = cmul<real>(re, im, state[i]);
    } else {
      state[i] = make_c<real>(0, 0);
    }
  }
}

// ---- gather / diagonal ------------------------------------------------------

template <typename real>
__global__ void sv_gather_kernel(const typename Cplx<real>::type* __restrict__ state,
                                 const uint64_t* __restrict__ idx, uint64_t count,
                                 double2* __restrict__ out) {
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  const auto a = state[idx[j]];
  out[j] = make_double2((double)a.x, (double)a.y);
}

template <typename real>
__global__ void dm_diag_kernel(const typename Cplx<real>::type* __restrict__ rho, int n,
                               double* __restrict__ probs) {
  const uint64_t dim = 1ull << n;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < dim;
       i += (uint64_t)gridDim.x * blockDim.x)
    probs[i] = (double)rho[i * dim + i].x;
}

// ---- marginal probabilities -------------------------------------------------

struct MarginalParams {
  int n;
  int zb;             // zone bits (6 for c64, 5 for c128)
  int m;              // measured bits
  int bits[24];       // measured bit positions, bits[0] = MSB of the key
  int n_meas_high;    // measured bits >= zb
  int meas_high_pos[24];  // ascending positions relative to zb
  int log2_iters;     // iterations per warp over unmeasured high bits
  uint64_t num_warps; // total warp tasks
};

// Each warp owns the 512-byte zone; each lane accumulates the amplitudes it
// streams into at most two private float64 sums (its key never changes while it
// iterates over the unmeasured high bits), then lanes that share a key are
// folded by shuffles and one atomicAdd per surviving lane hits probs[key].
template <typename real>
__global__ void __launch_bounds__(256)
    sv_marginal_kernel(const typename Cplx<real>::type* __restrict__ state,
                       const __grid_constant__ MarginalParams p, double* __restrict__ probs) {
  using C = typename Cplx<real>::type;
  constexpr bool kVec = sizeof(real) == 4;
  const int lane = threadIdx.x & 31;
  const uint64_t task = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (task >= p.num_warps) return;
  // task enumerates all high bits except the lowest `log2_iters` unmeasured ones.
  const uint64_t iters = 1ull << p.log2_iters;
  double a0 = 0.0, a1 = 0.0;
  // High index (bits >= zb) = insert_zero_bits(unmeasured combo, measured positions) | measured value.
  // task = (meas_value << (unmeasured_high - log2_iters)) | upper part of the unmeasured combo.
  const int n_high = p.n - p.zb;
  const int n_unmeas = n_high - p.n_meas_high;
  const int up_bits = n_unmeas - p.log2_iters;
  const uint64_t up = task & ((1ull << up_bits) - 1ull);
  const uint64_t mv = task >> up_bits;
  uint64_t meas_dep = 0;
  for (int i = 0; i < p.n_meas_high; ++i) meas_dep |= ((mv >> i) & 1ull) << p.meas_high_pos[i];
  uint64_t first_index = 0;
  for (uint64_t it = 0; it < iters; ++it) {
    const uint64_t combo = (up << p.log2_iters) | it;
    const uint64_t high = insert_zero_bits(combo, p.meas_high_pos, p.n_meas_high) | meas_dep;
    const uint64_t idx = (high << p.zb) | (uint64_t)(kVec ? (lane << 1) : lane);
    if (it == 0) first_index = idx;
    if constexpr (kVec) {
      const float4 v = *reinterpret_cast<const float4*>(state + idx);
      a0 += (double)(v.x * v.x + v.y * v.y);
      a1 += (double)(v.z * v.z + v.w * v.w);
    } else {
      const C v = state[idx];
      a0 += abs2<real>(v);
    }
  }
  // Fold lanes whose keys coincide.
  uint64_t meas_mask = 0;
  for (int q = 0; q < p.m; ++q) meas_mask |= 1ull << p.bits[q];
  bool alive = true;
  if constexpr (kVec) {
    if (!(meas_mask & 1ull)) {
      a0 += a1;
      a1 = 0.0;
    }
  }
  const int vb = kVec ? 1 : 0;
  for (int lb = 0; lb < 5; ++lb) {
    if (!((meas_mask >> (lb + vb)) & 1ull)) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, 1 << lb);
      if constexpr (kVec) a1 += __shfl_xor_sync(0xffffffffu, a1, 1 << lb);
      if ((lane >> lb) & 1) alive = false;
    }
  }
  if (alive) {
    const uint64_t k0 = extract_bits_msb_first(first_index, p.bits, p.m);
    atomicAdd(&probs[k0], a0);
    if constexpr (kVec) {
      if (meas_mask & 1ull) {
        const uint64_t k1 = extract_bits_msb_first(first_index | 1ull, p.bits, p.m);
        atomicAdd(&probs[k1], a1);
      }
    }
  }
}

// Small-state fallback (n < zone): one thread per amplitude, global atomics.
template <typename real>
__global__ void sv_marginal_small_kernel(const typename Cplx<real>::type* __restrict__ state,
                                         const __grid_constant__ MarginalParams p,
                                         double* __restrict__ probs) {
  const uint64_t total = 1ull << p.n;
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  atomicAdd(&probs[extract_bits_msb_first(i, p.bits, p.m)], abs2<real>(state[i]));
}

// Marginal of an explicit probability vector (density-matrix diagonals, n <= 20).
__global__ void probs_marginal_kernel(const double* __restrict__ probs,
                                      const __grid_constant__ MarginalParams p,
                                      double* __restrict__ out) {
  const uint64_t total = 1ull << p.n;
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  atomicAdd(&out[extract_bits_msb_first(i, p.bits, p.m)], probs[i]);
}

// ---- prefix scan (single CTA) -----------------------------------------------

// out[i] = in[0] + ... + in[i]; count arbitrary; 1024 threads, contiguous
// chunk per thread, block scan of the chunk totals.
__global__ void __launch_bounds__(1024)
    scan_inclusive_kernel(const double* __restrict__ in, double* __restrict__ out, uint64_t count) {
  __shared__ double warp_tot[32];
  __shared__ double carry_s;
  const int tid = threadIdx.x;
  const int lane = tid & 31, wid = tid >> 5;
  const uint64_t per = (count + 1023) / 1024;
  const uint64_t lo = std::min<uint64_t>(count, per * tid);
  const uint64_t hi = std::min<uint64_t>(count, lo + per);
  double t = 0.0;
  for (uint64_t i = lo; i < hi; ++i) t += in[i];
  double inc = warp_inclusive_scan(t, lane);
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    double w = warp_tot[lane];
    w = warp_inclusive_scan(w, lane);
    warp_tot[lane] = w;
  }
  __syncthreads();
  double prefix = inc - t + (wid > 0 ? warp_tot[wid - 1] : 0.0);
  (void)carry_s;
  for (uint64_t i = lo; i < hi; ++i) {
    prefix += in[i];
    out[i] = prefix;
  }
}

// ---- hierarchical sampler ---------------------------------------------------

// chunk_sums[c] = sum_{i in chunk c} |psi_i|^2 ; one warp per chunk.
template <typename real>
__global__ void __launch_bounds__(256)
    sv_chunk_sums_kernel(const typename Cplx<real>::type* __restrict__ state, uint64_t total,
                         int chunk_log2, double* __restrict__ chunk_sums) {
  using C = typename Cplx<real>::type;
  const int lane = threadIdx.x & 31;
  const uint64_t nchunks = total >> chunk_log2;
  const uint64_t chunk_elems = 1ull << chunk_log2;
  for (uint64_t c = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < nchunks;
       c += (uint64_t)gridDim.x * (blockDim.x >> 5)) {
    const C* base = state + (c << chunk_log2);
    double acc = 0.0;
    // lane owns 4 consecutive amplitudes (fewer for tiny chunks)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const uint64_t i = (uint64_t)lane * 4 + e;
      if (i < chunk_elems) acc += abs2<real>(base[i]);
    }
    acc = warp_sum(acc);
    if (lane == 0) chunk_sums[c] = acc;
  }
}

// block_sums[b] = sum of the block's chunk sums (warp per block).
__global__ void __launch_bounds__(256)
    block_sums_kernel(const double* __restrict__ chunk_sums, uint64_t nblocks, int cpb_log2,
                      double* __restrict__ block_sums) {
  const int lane = threadIdx.x & 31;
  const uint64_t cpb = 1ull << cpb_log2;
  for (uint64_t b = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < nblocks;
       b += (uint64_t)gridDim.x * (blockDim.x >> 5)) {
    const double* base = chunk_sums + (b << cpb_log2);
    double acc = 0.0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const uint64_t i = (uint64_t)lane * 4 + e;
      if (i < cpb) acc += base[i];
    }
    acc = warp_sum(acc);
    if (lane == 0) block_sums[b] = acc;
  }
}

// Finds, within 4-per-lane values v[0..3] (lane-major order), the first global
// position whose inclusive prefix exceeds `target`.  Returns position or -1, and
// the exclusive prefix at that position through *before.
__device__ __forceinline__ int warp_search4(const double (&v)[4], int valid, double target,
                                            int lane, double* before) {
  const double lane_tot = v[0] + v[1] + v[2] + v[3];
  const double inc = warp_inclusive_scan(lane_tot, lane);
  const double exc = inc - lane_tot;
  int pos = -1;
  double bef = 0.0;
  double run = exc;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int gi = lane * 4 + e;
    if (pos < 0 && gi < valid && run + v[e] > target) {
      pos = gi;
      bef = run;
    }
    run += v[e];
  }
  // first lane that found something wins
  const unsigned ball = __ballot_sync(0xffffffffu, pos >= 0);
  if (ball == 0) return -1;
  const int src = __ffs(ball) - 1;
  pos = __shfl_sync(0xffffffffu, pos, src);
  bef = __shfl_sync(0xffffffffu, bef, src);
  *before = bef;
  return pos;
}

// Last position with a strictly positive value, or -1.
__device__ __forceinline__ int warp_last_positive4(const double (&v)[4], int valid, int lane) {
  int pos = -1;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int gi = lane * 4 + e;
    if (gi < valid && v[e] > 0.0) pos = gi;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) pos = max(pos, __shfl_xor_sync(0xffffffffu, pos, o));
  return pos;
}

template <typename real>
__global__ void __launch_bounds__(256)
    sv_sample_resolve_kernel(const typename Cplx<real>::type* __restrict__ state, uint64_t total,
                             int chunk_log2, int cpb_log2, const double* __restrict__ chunk_sums,
                             const double* __restrict__ block_cum, uint64_t nblocks,
                             const double* __restrict__ uniforms, uint64_t reps,
                             uint64_t* __restrict__ out) {
  using C = typename Cplx<real>::type;
  const int lane = threadIdx.x & 31;
  const double grand = block_cum[nblocks - 1];
  const int chunk_elems = 1 << chunk_log2;
  const int cpb = 1 << cpb_log2;
  for (uint64_t j = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); j < reps;
       j += (uint64_t)gridDim.x * (blockDim.x >> 5)) {
    double target = uniforms[j] * grand;
    // smallest b with block_cum[b] > target
    uint64_t lo = 0, hi = nblocks;
    while (lo < hi) {
      const uint64_t mid = (lo + hi) >> 1;
      if (block_cum[mid] > target) {
        hi = mid;
      } else {
        lo = mid + 1;
      }
    }
    uint64_t b = lo;
    if (b >= nblocks) b = nblocks - 1;
    target -= (b > 0 ? block_cum[b - 1] : 0.0);
    // chunk level
    double v[4];
    const double* cs = chunk_sums + (b << cpb_log2);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int gi = lane * 4 + e;
      v[e] = gi < cpb ? cs[gi] : 0.0;
    }
    double before = 0.0;
    int c = warp_search4(v, cpb, target, lane, &before);
    if (c < 0) {
      c = warp_last_positive4(v, cpb, lane);
      if (c < 0) c = cpb - 1;
      before = target;  // forces the in-chunk fallback below
      // exclusive prefix unknown: make the in-chunk search pick its last positive entry
      target = 1e300;
    } else {
      target -= before;
    }
    const uint64_t chunk = (b << cpb_log2) + (uint64_t)c;
    const C* base = state + (chunk << chunk_log2);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int gi = lane * 4 + e;
      v[e] = gi < chunk_elems ? abs2<real>(base[gi]) : 0.0;
    }
    int pos = warp_search4(v, chunk_elems, target, lane, &before);
    if (pos < 0) {
      pos = warp_last_positive4(v, chunk_elems, lane);
      if (pos < 0) pos = chunk_elems - 1;
    }
    if (lane == 0) out[j] = (chunk << chunk_log2) + (uint64_t)pos;
  }
}

// Thread per sample: searchsorted(cum, u*total, 'right') over a small table.
__global__ void cdf_search_kernel(const double* __restrict__ cum, uint64_t count,
                                  const double* __restrict__ uniforms, uint64_t reps,
                                  uint64_t* __restrict__ out) {
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= reps) return;
  const double target = uniforms[j] * cum[count - 1];
  uint64_t lo = 0, hi = count;
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    if (cum[mid] > target) {
      hi = mid;
    } else {
      lo = mid + 1;
    }
  }
  if (lo >= count) {
    // rounding pushed the target to the very end: last entry with positive mass
    lo = count - 1;
    while (lo > 0 && !(cum[lo] > cum[lo - 1])) --lo;
  }
  out[j] = lo;
}

struct UnpackParams {
  int m;
  int bits[64];
};

__global__ void unpack_bits_kernel(const uint64_t* __restrict__ idx, uint64_t total_out,
                                   const __grid_constant__ UnpackParams p,
                                   uint8_t* __restrict__ out) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total_out) return;
  const uint64_t j = t / (uint64_t)p.m;
  const int q = (int)(t - j * (uint64_t)p.m);
  out[t] = (uint8_t)((idx[j] >> p.bits[q]) & 1ull);
}

// ---- Pauli expectation ------------------------------------------------------

// partial[b] = sum_i sign(i) * conj(psi[i ^ x]) * psi[i]   (complex, 2 doubles)
template <typename real>
__global__ void __launch_bounds__(256)
    sv_pauli_partial_kernel(const typename Cplx<real>::type* __restrict__ state, uint64_t total,
                            uint64_t xmask, uint64_t zmask, double* __restrict__ partial) {
  using C = typename Cplx<real>::type;
  double ar = 0.0, ai = 0.0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const C a = state[i];
    const C b = state[i ^ xmask];
    // conj(b) * a
    double re = (double)b.x * (double)a.x + (double)b.y * (double)a.y;
    double im = (double)b.x * (double)a.y - (double)b.y * (double)a.x;
    if (__popcll(i & zmask) & 1) {
      re = -re;
      im = -im;
    }
    ar += re;
    ai += im;
  }
  __shared__ double sm[16];
  ar = warp_sum(ar);
  ai = warp_sum(ai);
  if ((threadIdx.x & 31) == 0) {
    sm[2 * (threadIdx.x >> 5)] = ar;
    sm[2 * (threadIdx.x >> 5) + 1] = ai;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tr = 0, ti = 0;
    for (int w = 0; w < 8; ++w) {
      tr += sm[2 * w];
      ti += sm[2 * w + 1];
    }
    partial[2 * blockIdx.x] = tr;
    partial[2 * blockIdx.x + 1] = ti;
  }
}

// Several Pauli strings that flip the SAME bits (one x mask, T <= 16 z masks) in ONE
// pass: conj(psi[i ^ x]) * psi[i] is formed once per amplitude, each string only adds
// its own sign.  A PauliSum of Z-type terms (the cost Hamiltonian of a QAOA sweep:
// x = 0 for every edge) costs one read of the state instead of one per term.
// partial[b * 2T + 2t] (+1) = re (im) of block b's sum for string t.
constexpr int kPauliMaxTerms = 16;
struct PauliMultiParams {
  uint64_t zmask[kPauliMaxTerms];
  int count;
};

// DIAG: x mask 0 (Z-type strings): the pair product is |psi[i]|^2, real.
template <typename real, bool DIAG>
__global__ void __launch_bounds__(256)
    sv_pauli_multi_partial_kernel(const typename Cplx<real>::type* __restrict__ state,
                                  uint64_t total, uint64_t xmask,
                                  const __grid_constant__ PauliMultiParams p,
                                  double* __restrict__ partial) {
  using C = typename Cplx<real>::type;
  double ar[kPauliMaxTerms], ai[kPauliMaxTerms];
#pragma unroll
  for (int t = 0; t < kPauliMaxTerms; ++t) ar[t] = ai[t] = 0.0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const C a = state[i];
    const C b = DIAG ? a : state[i ^ xmask];
    const double re = (double)b.x * (double)a.x + (double)b.y * (double)a.y;
    const double im = DIAG ? 0.0 : (double)b.x * (double)a.y - (double)b.y * (double)a.x;
#pragma unroll
    for (int t = 0; t < kPauliMaxTerms; ++t) {
      if (t < p.count) {
        const bool neg = __popcll(i & p.zmask[t]) & 1;
        ar[t] += neg ? -re : re;
        if constexpr (!DIAG) ai[t] += neg ? -im : im;
      }
    }
  }
  __shared__ double sm[8][2 * kPauliMaxTerms];
#pragma unroll
  for (int t = 0; t < kPauliMaxTerms; ++t) {
    if (t < p.count) {
      const double r = warp_sum(ar[t]);
      const double m = warp_sum(ai[t]);
      if ((threadIdx.x & 31) == 0) {
        sm[threadIdx.x >> 5][2 * t] = r;
        sm[threadIdx.x >> 5][2 * t + 1] = m;
      }
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < 2 * p.count) {
    double v = 0;
    for (int w = 0; w < 8; ++w) v += sm[w][threadIdx.x];
    partial[(size_t)blockIdx.x * 2 * p.count + threadIdx.x] = v;
  }
}

// The same sums with the sign work amortised.  A thread takes a run of 8 amplitudes
// (index bits 0-2) per step: the 8 signed sums of the run over its low bits are the
// Walsh-Hadamard transform W of its 8 pair products (24 additions for all 8), so a
// string with Z bits z reads W[z & 7].  The rest of the sign splits by index bits:
// run r = k * vt + g with vt (a power of two) "virtual threads" g — the part of z over
// g's bits is constant while a thread walks k and is applied once per walk, the part
// over k's bits is the same for every thread and comes from a table built once per
// CTA (flip[k][t] = sign bit to XOR into the addend).  Per run and string that leaves
// one shared-memory read of W (own column: no synchronisation), one XOR and one
// float64 addition.  A CTA walks virtual CTAs and keeps running column sums, so
// partial[] has gridDim.x rows.
constexpr int kPauliMaxK = 128;
template <typename real, bool DIAG>
__global__ void __launch_bounds__(256)
    sv_pauli_multi_run_kernel(const typename Cplx<real>::type* __restrict__ state, uint64_t vt,
                              int k_count, uint64_t xmask,
                              const __grid_constant__ PauliMultiParams p,
                              double* __restrict__ partial) {
  using C = typename Cplx<real>::type;
  constexpr int T = kPauliMaxTerms;
  __shared__ double w_re[8][256];
  __shared__ double w_im[DIAG ? 1 : 8][DIAG ? 1 : 256];
  __shared__ __align__(16) uint32_t flip[kPauliMaxK][T];
  __shared__ double sm[8][2 * T];
  const int tid = threadIdx.x;
  for (int idx = tid; idx < k_count * T; idx += 256) {
    const int k = idx / T, t = idx % T;
    flip[k][t] = (uint32_t)(__popcll((((uint64_t)k * vt) << 3) & p.zmask[t]) & 1) << 31;
  }
  __syncthreads();
  const uint64_t xhigh = xmask & ~7ull;
  const int xlow = (int)(xmask & 7ull);
  int w_at[T];  // this thread's element of W[z_t & 7]
#pragma unroll
  for (int t = 0; t < T; ++t) w_at[t] = (int)(p.zmask[t] & 7ull) * 256 + tid;
  double column_total = 0.0;  // threads < 2 * count: this CTA's sum of one output column
  for (uint64_t vb = blockIdx.x; vb < (vt >> 8); vb += gridDim.x) {
    double ar[T], ai[T];
#pragma unroll
    for (int t = 0; t < T; ++t) ar[t] = ai[t] = 0.0;
    const uint64_t g = (vb << 8) | (uint64_t)tid;
    for (int k = 0; k < k_count; ++k) {
      const uint64_t i0 = ((uint64_t)k * vt + g) << 3;
      C a[8];
      if constexpr (sizeof(real) == 4) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const float4 v = *reinterpret_cast<const float4*>(state + i0 + j);
          a[j] = make_float2(v.x, v.y);
          a[j + 1] = make_float2(v.z, v.w);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = state[i0 + j];
      }
      double wr[8], wi[8];
      if constexpr (DIAG) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          wr[j] = (double)a[j].x * (double)a[j].x + (double)a[j].y * (double)a[j].y;
          wi[j] = 0.0;
        }
      } else {
        const C* other = state + (i0 ^ xhigh);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const C b = other[j ^ xlow];
          wr[j] = (double)b.x * (double)a[j].x + (double)b.y * (double)a[j].y;
          wi[j] = (double)b.x * (double)a[j].y - (double)b.y * (double)a[j].x;
        }
      }
#pragma unroll
      for (int h = 1; h < 8; h <<= 1)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (!(j & h)) {
            const double x = wr[j], y = wr[j | h];
            wr[j] = x + y;
            wr[j | h] = x - y;
            if constexpr (!DIAG) {
              const double u = wi[j], v = wi[j | h];
              wi[j] = u + v;
              wi[j | h] = u - v;
            }
          }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        w_re[j][tid] = wr[j];
        if constexpr (!DIAG) w_im[j][tid] = wi[j];
      }
      uint32_t f[T];
#pragma unroll
      for (int q = 0; q < T; q += 4) {
        const uint4 v = *reinterpret_cast<const uint4*>(&flip[k][q]);
        f[q] = v.x, f[q + 1] = v.y, f[q + 2] = v.z, f[q + 3] = v.w;
      }
#pragma unroll
      for (int t = 0; t < T; ++t) {
        if (t < p.count) {
          const double re = (&w_re[0][0])[w_at[t]];
          ar[t] += __hiloint2double(__double2hiint(re) ^ (int)f[t], __double2loint(re));
          if constexpr (!DIAG) {
            const double im = (&w_im[0][0])[w_at[t]];
            ai[t] += __hiloint2double(__double2hiint(im) ^ (int)f[t], __double2loint(im));
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < T; ++t) {
      if (t < p.count) {
        const bool neg = __popcll((g << 3) & p.zmask[t]) & 1;
        const double r = warp_sum(neg ? -ar[t] : ar[t]);
        const double m = DIAG ? 0.0 : warp_sum(neg ? -ai[t] : ai[t]);
        if ((tid & 31) == 0) {
          sm[tid >> 5][2 * t] = r;
          sm[tid >> 5][2 * t + 1] = m;
        }
      }
    }
    __syncthreads();
    if (tid < 2 * p.count) {
      double v = 0;
      for (int w = 0; w < 8; ++w) v += sm[w][tid];
      column_total += v;
    }
    __syncthreads();
  }
  if (tid < 2 * p.count) partial[(size_t)blockIdx.x * 2 * p.count + tid] = column_total;
}

// Launch shape of the run kernel: run = k * vt + virtual thread, at most kPauliMaxK values
// of k, vt >= 2^18 virtual threads for big states (enough virtual CTAs for every SM).
struct PauliRunPlan {
  bool by_runs;
  uint64_t vt;
  int k_count;
  unsigned blocks;
};
static PauliRunPlan pauli_run_plan(int n_qubits, bool runs_enabled) {
  PauliRunPlan r;
  const uint64_t total = 1ull << n_qubits;
  r.by_runs = runs_enabled && n_qubits >= 12;
  const uint64_t num_runs = total >> 3;
  r.vt = r.by_runs ? std::min<uint64_t>(num_runs, std::max<uint64_t>(1ull << 18, num_runs / kPauliMaxK)) : 1;
  r.k_count = (int)(num_runs / r.vt);
  r.blocks = r.by_runs ? (unsigned)std::min<uint64_t>(r.vt >> 8, 148ull * 4) : 0;
  return r;
}

// out[j] = sum_b partial[b * width + j]
__global__ void __launch_bounds__(256)
    column_sum_kernel(const double* __restrict__ partial, uint64_t blocks, int width,
                      double* __restrict__ out) {
  __shared__ double sm[256];
  for (int j = 0; j < width; ++j) {
    double v = 0;
    for (uint64_t b = threadIdx.x; b < blocks; b += blockDim.x) v += partial[b * width + j];
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if ((int)threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[j] = sm[0];
    __syncthreads();
  }
}

// ---- reduced density matrix ---------------------------------------------------

// rho[a][b] = sum_rest psi[a, rest] conj(psi[b, rest]) over the M kept bits
// (qis/states.py:623-693 density_matrix_from_state_vector with `indices`).  A
// thread walks `rest` values grid-stride, holds the D = 2^M amplitudes of one
// rest value in registers and accumulates R rows of the outer product (row
// tile = blockIdx.y, so wide reductions re-read the state D/R times); partial
// sums are folded per warp in float64 and added to out[D*D*2] with one atomic
// per entry per warp.  sorted_pos = kept bit positions ascending; index a uses
// them LSB first (the host permutes to the caller's qubit order).
struct RdmParams {
  int pos[6];
};

template <typename real, int M, int R>
__global__ void __launch_bounds__(256)
    sv_reduced_dm_kernel(const typename Cplx<real>::type* __restrict__ state, uint64_t rest_total,
                         const __grid_constant__ RdmParams p, double* __restrict__ out) {
  using C = typename Cplx<real>::type;
  constexpr int D = 1 << M;
  const int row0 = blockIdx.y * R;
  // per-thread partial sums in the state's precision (a thread sees a few hundred
  // terms); the cross-thread fold is float64
  real acc_re[R][D], acc_im[R][D];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int b = 0; b < D; ++b) acc_re[r][b] = acc_im[r][b] = 0;
  uint64_t row_off[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    row_off[r] = 0;
#pragma unroll
    for (int b = 0; b < M; ++b)
      if (((row0 + r) >> b) & 1) row_off[r] += 1ull << p.pos[b];
  }
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < rest_total;
       g += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t base = insert_zero_bits(g, p.pos, M);
    C xa[R];
#pragma unroll
    for (int r = 0; r < R; ++r) xa[r] = state[base + row_off[r]];
#pragma unroll
    for (int b = 0; b < D; ++b) {
      uint64_t off = 0;
#pragma unroll
      for (int q = 0; q < M; ++q)
        if ((b >> q) & 1) off += 1ull << p.pos[q];
      const C xb = state[base + off];  // the row loads above brought the lines in
#pragma unroll
      for (int r = 0; r < R; ++r) {
        // xa * conj(xb)
        acc_re[r][b] = fma(xa[r].x, xb.x, fma(xa[r].y, xb.y, acc_re[r][b]));
        acc_im[r][b] = fma(xa[r].y, xb.x, fma(-xa[r].x, xb.y, acc_im[r][b]));
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int b = 0; b < D; ++b) {
      const double re = warp_sum((double)acc_re[r][b]);
      const double im = warp_sum((double)acc_im[r][b]);
      if ((threadIdx.x & 31) == 0) {
        atomicAdd(&out[2 * ((row0 + r) * D + b)], re);
        atomicAdd(&out[2 * ((row0 + r) * D + b) + 1], im);
      }
    }
}

template <typename real, int M, int R>
static int reduced_dm_launch(const void* state, int n, const RdmParams& p, double* out_dev,
                             cudaStream_t s) {
  using C = typename Cplx<real>::type;
  constexpr int D = 1 << M;
  const uint64_t rest_total = 1ull << (n - M);
  const unsigned bx = (unsigned)std::max<uint64_t>(
      1, std::min<uint64_t>((rest_total + 255) / 256, 148ull * 4));
  const dim3 grid(bx, D / R);
  sv_reduced_dm_kernel<real, M, R>
      <<<grid, 256, 0, s>>>(reinterpret_cast<const C*>(state), rest_total, p, out_dev);
  B2Q_LAUNCH_CHECK("sv_reduced_dm_kernel");
  return B2Q_OK;
}

template <typename real>
static int reduced_dm_t(const void* state, int n, int m, const RdmParams& p, double* out_dev,
                        cudaStream_t s) {
  switch (m) {
    case 1: return reduced_dm_launch<real, 1, 2>(state, n, p, out_dev, s);
    case 2: return reduced_dm_launch<real, 2, 4>(state, n, p, out_dev, s);
    case 3: return reduced_dm_launch<real, 3, sizeof(real) == 4 ? 4 : 2>(state, n, p, out_dev, s);
    case 4: return reduced_dm_launch<real, 4, sizeof(real) == 4 ? 2 : 1>(state, n, p, out_dev, s);
    default: return reduced_dm_launch<real, 5, 1>(state, n, p, out_dev, s);
  }
}

// Gram form of the same sum for M = 3..5 kept bits and states of >= 2^11
// amplitudes: ONE read of the state.  A CTA walks tiles of 2^11 amplitudes — the
// kept bits plus the lowest free bits, so a tile is made of long contiguous runs —,
// stages a tile in shared memory as X[rest][a] and accumulates rho += X^T conj(X)
// with 4 x 4 register blocks (thread = one block of rho x one slice of the tile's
// rest values; complex64: two packed FFMA2 per complex MAC, partial sums in the
// state's precision over one tile, folded into float64 registers per tile).  The
// next tile's loads are in flight while the current one is multiplied.  CTA sums
// go to partial[cta][2 D D] and rdm_fold_kernel adds them in a fixed order: the
// result does not depend on scheduling.
constexpr int kRdmTileBits = 11;
struct RdmGramParams {
  int tile_pos[kRdmTileBits];  // ascending: the kept bits and the lowest free bits
  int kept_rank[5];            // kept bit q (ascending position) = tile bit kept_rank[q]
};

// Element e (11 bits, tile bits ascending) of a tile: offset of its amplitude from the
// tile's base, and its slot in the staged tile X[rest][a] (a = the kept bits, LSB first).
B2Q_HD uint64_t rdm_tile_offset(const RdmGramParams& p, int e) {
  uint64_t off = 0;
  for (int j = 0; j < kRdmTileBits; ++j) off |= (uint64_t)((e >> j) & 1) << p.tile_pos[j];
  return off;
}
B2Q_HD int rdm_tile_slot(const RdmGramParams& p, int m, int e) {
  int a = 0, kept = 0;
  for (int q = 0; q < m; ++q) {
    a |= ((e >> p.kept_rank[q]) & 1) << q;
    kept |= 1 << p.kept_rank[q];
  }
  int rest = 0, f = 0;
  for (int j = 0; j < kRdmTileBits; ++j)
    if (!((kept >> j) & 1)) rest |= ((e >> j) & 1) << f++;
  return (rest << m) | a;
}

// Tile bits of a reduction over the ascending kept bits `sorted`: the kept bits and the
// lowest free bits (a free bit is taken while there is room for the kept bits to come).
static RdmGramParams rdm_gram_plan(const int* sorted, int m) {
  RdmGramParams gp;
  int count = 0;
  for (int b = 0, kept = 0; count < kRdmTileBits; ++b) {
    const bool is_kept = std::find(sorted, sorted + m, b) != sorted + m;
    if (is_kept) ++kept;
    if (is_kept || count - kept < kRdmTileBits - m) gp.tile_pos[count++] = b;
  }
  for (int q = 0; q < 5; ++q)
    gp.kept_rank[q] =
        q < m ? (int)(std::find(gp.tile_pos, gp.tile_pos + kRdmTileBits, sorted[q]) - gp.tile_pos) : 0;
  return gp;
}

template <typename real>
__device__ __forceinline__ void rdm_mac(typename Cplx<real>::type& acc,
                                        const typename Cplx<real>::type& xa,
                                        const typename Cplx<real>::type& xb);
template <>
__device__ __forceinline__ void rdm_mac<float>(float2& acc, const float2& xa, const float2& xb) {
  // acc += xa * conj(xb)
  acc = __ffma2_rn(make_float2(xb.x, xb.x), xa, acc);
  acc = __ffma2_rn(make_float2(xb.y, -xb.y), make_float2(xa.y, xa.x), acc);
}
template <>
__device__ __forceinline__ void rdm_mac<double>(double2& acc, const double2& xa,
                                                const double2& xb) {
  acc.x = fma(xa.x, xb.x, fma(xa.y, xb.y, acc.x));
  acc.y = fma(xa.y, xb.x, fma(-xa.x, xb.y, acc.y));
}

template <typename real, int M>
__global__ void __launch_bounds__(256)
    sv_reduced_dm_gram_kernel(const typename Cplx<real>::type* __restrict__ state,
                              uint64_t num_tiles, const __grid_constant__ RdmGramParams p,
                              double* __restrict__ partial) {
  using C = typename Cplx<real>::type;
  constexpr int D = 1 << M;
  constexpr int E = 1 << kRdmTileBits;  // amplitudes per tile
  constexpr int G = E / D;              // rest values per tile
  constexpr int NBR = D / 4;            // 4 x 4 blocks per side of rho
  constexpr int NB = NBR * NBR;
  constexpr int S = 256 / NB;           // slices of the rest values
  constexpr int kLoads = E / 256;
  static_assert(M >= 3 && M <= 5 && G % S == 0, "tile shape");
  static_assert(sizeof(C) * E >= sizeof(double) * 2 * D * D, "fold buffer aliases the tile");
  __shared__ __align__(16) C tile[E];
  __shared__ uint64_t k_off[kLoads];
  __shared__ int k_slot[kLoads];

  const int t = threadIdx.x;
  // element e = t | (k << 8) of a tile: its offset in the state and its slot X[rest][a]
  auto offset_of = [&](int e) { return rdm_tile_offset(p, e); };
  auto slot_of = [&](int e) { return rdm_tile_slot(p, M, e); };
  const uint64_t t_off = offset_of(t);
  const int t_slot = slot_of(t);
  if (t < kLoads) {
    k_off[t] = offset_of(t << 8);
    k_slot[t] = slot_of(t << 8);
  }
  __syncthreads();

  const int blk = t % NB, slice = t / NB;
  const int ra = 4 * (blk / NBR), rb = 4 * (blk % NBR);
  double acc_re[4][4], acc_im[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc_re[r][c] = acc_im[r][c] = 0.0;

  C v[kLoads];
  uint64_t tl = blockIdx.x;
  if (tl < num_tiles) {
    const C* src = state + insert_zero_bits(tl, p.tile_pos, kRdmTileBits) + t_off;
#pragma unroll
    for (int k = 0; k < kLoads; ++k) v[k] = src[k_off[k]];
  }
  for (; tl < num_tiles; tl += gridDim.x) {
    __syncthreads();  // the previous tile has been consumed
#pragma unroll
    for (int k = 0; k < kLoads; ++k) tile[t_slot | k_slot[k]] = v[k];
    __syncthreads();
    const uint64_t next = tl + gridDim.x;
    if (next < num_tiles) {
      const C* src = state + insert_zero_bits(next, p.tile_pos, kRdmTileBits) + t_off;
#pragma unroll
      for (int k = 0; k < kLoads; ++k) v[k] = src[k_off[k]];
    }
    C part[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) part[r][c] = make_c<real>(0, 0);
#pragma unroll 4
    for (int g = slice; g < G; g += S) {
      const C* row = tile + g * D;
      C xa[4], xb[4];
      if constexpr (sizeof(real) == 4) {
        const float4 a01 = *reinterpret_cast<const float4*>(row + ra);
        const float4 a23 = *reinterpret_cast<const float4*>(row + ra + 2);
        const float4 b01 = *reinterpret_cast<const float4*>(row + rb);
        const float4 b23 = *reinterpret_cast<const float4*>(row + rb + 2);
        xa[0] = make_float2(a01.x, a01.y), xa[1] = make_float2(a01.z, a01.w);
        xa[2] = make_float2(a23.x, a23.y), xa[3] = make_float2(a23.z, a23.w);
        xb[0] = make_float2(b01.x, b01.y), xb[1] = make_float2(b01.z, b01.w);
        xb[2] = make_float2(b23.x, b23.y), xb[3] = make_float2(b23.z, b23.w);
      } else {
#pragma unroll
        for (int r = 0; r < 4; ++r) xa[r] = row[ra + r], xb[r] = row[rb + r];
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) rdm_mac<real>(part[r][c], xa[r], xb[c]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        acc_re[r][c] += (double)part[r][c].x;
        acc_im[r][c] += (double)part[r][c].y;
      }
  }
  // fold the slices in a fixed order (the tile's memory is free now)
  __syncthreads();
  double* red = reinterpret_cast<double*>(tile);
  for (int i = t; i < 2 * D * D; i += 256) red[i] = 0.0;
  __syncthreads();
  for (int s = 0; s < S; ++s) {
    if (slice == s) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          red[2 * ((ra + r) * D + rb + c)] += acc_re[r][c];
          red[2 * ((ra + r) * D + rb + c) + 1] += acc_im[r][c];
        }
    }
    __syncthreads();
  }
  for (int i = t; i < 2 * D * D; i += 256) partial[(size_t)blockIdx.x * 2 * D * D + i] = red[i];
}

// out[i] = sum over CTAs of partial[cta][i], in CTA order
__global__ void __launch_bounds__(256)
    rdm_fold_kernel(const double* __restrict__ partial, int ctas, int width,
                    double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= width) return;
  double v = 0.0;
  for (int b = 0; b < ctas; ++b) v += partial[(size_t)b * width + i];
  out[i] = v;
}

// Most CTAs of the Gram kernel that can be resident at once (its loop is persistent).
template <typename real, int M>
static int rdm_gram_grid(uint64_t num_tiles) {
  int dev = 0, sms = 148, per_sm = 1;
  if (cudaGetDevice(&dev) == cudaSuccess)
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sv_reduced_dm_gram_kernel<real, M>,
                                                    256, 0) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  return (int)std::min<uint64_t>(num_tiles, (uint64_t)sms * std::min(per_sm, 2));
}

template <typename real, int M>
static int reduced_dm_gram_launch(const void* state, int n, const RdmGramParams& p, int ctas,
                                  double* partial, double* out_dev, cudaStream_t s) {
  using C = typename Cplx<real>::type;
  constexpr int width = 2 << (2 * M);
  sv_reduced_dm_gram_kernel<real, M><<<ctas, 256, 0, s>>>(
      reinterpret_cast<const C*>(state), 1ull << (n - kRdmTileBits), p, partial);
  B2Q_LAUNCH_CHECK("sv_reduced_dm_gram_kernel");
  rdm_fold_kernel<<<(width + 255) / 256, 256, 0, s>>>(partial, ctas, width, out_dev);
  B2Q_LAUNCH_CHECK("rdm_fold_kernel");
  return B2Q_OK;
}

// tr(rho P) = sum_i rho[i, i ^ x] * sign(i): only 2^n of rho's 4^n entries are read.
// partial[b] = the block's (re, im) sum.
template <typename real>
__global__ void __launch_bounds__(256)
    dm_pauli_partial_kernel(const typename Cplx<real>::type* __restrict__ rho, int n,
                            uint64_t xmask, uint64_t zmask, double* __restrict__ partial) {
  using C = typename Cplx<real>::type;
  const uint64_t dim = 1ull << n;
  double ar = 0.0, ai = 0.0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < dim;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const C a = rho[(i << n) | (i ^ xmask)];
    const bool neg = __popcll(i & zmask) & 1;
    ar += neg ? -(double)a.x : (double)a.x;
    ai += neg ? -(double)a.y : (double)a.y;
  }
  __shared__ double sm[16];
  ar = warp_sum(ar);
  ai = warp_sum(ai);
  if ((threadIdx.x & 31) == 0) {
    sm[2 * (threadIdx.x >> 5)] = ar;
    sm[2 * (threadIdx.x >> 5) + 1] = ai;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tr = 0, ti = 0;
    for (int w = 0; w < 8; ++w) {
      tr += sm[2 * w];
      ti += sm[2 * w + 1];
    }
    partial[2 * blockIdx.x] = tr;
    partial[2 * blockIdx.x + 1] = ti;
  }
}

// ---- dist pack / unpack -----------------------------------------------------

struct PackParams {
  int n_local;
  int g;
  int pos[8];  // ascending local bit positions being exchanged
  int rank_of[8];  // pos[i] <-> bit rank_of[i] of the segment index
};

// One thread per amplitude of the shard.  dir=0: shard -> packed; 1: packed -> shard.
template <typename real, int DIR>
__global__ void __launch_bounds__(256)
    dist_pack_kernel(typename Cplx<real>::type* __restrict__ shard,
                     typename Cplx<real>::type* __restrict__ packed,
                     const __grid_constant__ PackParams p) {
  const uint64_t total = 1ull << p.n_local;
  const int rest = p.n_local - p.g;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    // t = (segment << rest) | r  indexes `packed`
    const uint64_t seg = t >> rest;
    const uint64_t r = t & ((1ull << rest) - 1ull);
    uint64_t idx = insert_zero_bits(r, p.pos, p.g);
    for (int i = 0; i < p.g; ++i) idx |= ((seg >> p.rank_of[i]) & 1ull) << p.pos[i];
    if (DIR == 0) {
      packed[t] = shard[idx];
    } else {
      shard[idx] = packed[t];
    }
  }
}

// Grid for grid-stride kernels: enough CTAs to cover the work, capped at 16
// resident waves of 256 threads on the 148 SMs.
inline unsigned stride_grid(uint64_t work_items, int threads) {
  const uint64_t b = (work_items + threads - 1) / threads;
  return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(b, 148ull * 8 * 16));
}

template <typename real>
int norm2_t(const void* state, int n, double* out_host, cudaStream_t s) {
  using C = typename Cplx<real>::type;
  const uint64_t total = 1ull << n;
  if (total < 2) {
    // single amplitude: read it back directly
    C h;
    B2Q_CUDA_CHECK(cudaMemcpyAsync(&h, state, sizeof(C), cudaMemcpyDeviceToHost, s));
    B2Q_CUDA_CHECK(cudaStreamSynchronize(s));
    *out_host = (double)h.x * (double)h.x + (double)h.y * (double)h.y;
    return B2Q_OK;
  }
  const unsigned blocks = stride_grid(total / V16<real>::kElems / kStreamUnroll, 256);
  double* partial = reinterpret_cast<double*>(workspace(sizeof(double) * (blocks + 1)));
  if (partial == nullptr) return B2Q_ERR_CUDA;
  sv_norm_partial_kernel<real><<<blocks, 256, 0, s>>>(reinterpret_cast<const C*>(state), total,
                                                      partial);
  B2Q_LAUNCH_CHECK("sv_norm_partial_kernel");
  final_sum_kernel<<<1, 256, 0, s>>>(partial, blocks, 1, partial + blocks);
  B2Q_LAUNCH_CHECK("final_sum_kernel");
  B2Q_CUDA_CHECK(
      cudaMemcpyAsync(out_host, partial + blocks, sizeof(double), cudaMemcpyDeviceToHost, s));
  B2Q_CUDA_CHECK(cudaStreamSynchronize(s));
  return B2Q_OK;
}

}  // namespace b2q

using namespace b2q;

#define B2Q_CHECK_DTYPE(dtype) \
  B2Q_REQUIRE((dtype) == B2Q_C64 || (dtype) == B2Q_C128, "bad dtype %d", (dtype))

extern "C" int b2q_sv_init_basis(void* state, int dtype, int n_qubits, uint64_t basis_index,
                                 void* stream) {
  B2Q_REQUIRE(state != nullptr, "null state");
  B2Q_CHECK_DTYPE(dtype);
  B2Q_REQUIRE(n_qubits >= 0 && n_qubits <= 40, "n_qubits out of range");
  const uint64_t total = 1ull << n_qubits;
  B2Q_REQUIRE(basis_index < total, "basis index out of range");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  B2Q_CUDA_CHECK(cudaMemsetAsync(state, 0, total * elem_bytes(dtype), s));
  if (dtype == B2Q_C64)
    set_one_kernel<float><<<1, 1, 0, s>>>(reinterpret_cast<float2*>(state), basis_index);
  else
    set_one_kernel<double><<<1, 1, 0, s>>>(reinterpret_cast<double2*>(state), basis_index);
  B2Q_LAUNCH_CHECK("set_one_kernel");
  return B2Q_OK;
}

static int scale_mask_common(void* state, int dtype, uint64_t total, uint64_t mask, uint64_t want,
                             double re, double im, cudaStream_t s) {
  if (dtype == B2Q_C64) {
    if (total < 2) {
      sv_scale_tiny_kernel<float><<<1, 32, 0, s>>>(reinterpret_cast<float2*>(state), total, mask,
                                                   want, (float)re, (float)im);
    } else {
      const unsigned blocks = stride_grid(total / 2 / kStreamUnroll, 256);
      sv_scale_mask_kernel<float><<<blocks, 256, 0, s>>>(reinterpret_cast<float2*>(state), total,
                                                         mask, want, (float)re, (float)im);
    }
  } else {
    const unsigned blocks = stride_grid(total / kStreamUnroll, 256);
    sv_scale_mask_kernel<double><<<blocks, 256, 0, s>>>(reinterpret_cast<double2*>(state), total,
                                                        mask, want, re, im);
  }
  B2Q_LAUNCH_CHECK("sv_scale_mask_kernel");
  return B2Q_OK;
}

extern "C" int b2q_sv_scale(void* state, int dtype, int n_qubits, double re, double im,
                            void* stream) {
  B2Q_REQUIRE(state != nullptr, "null state");
  B2Q_CHECK_DTYPE(dtype);
  return scale_mask_common(state, dtype, 1ull << n_qubits, 0, 0, re, im,
                           reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int b2q_sv_norm2(const void* state, int dtype, int n_qubits, double* out,
                            void* stream) {
  B2Q_REQUIRE(state != nullptr && out != nullptr, "null argument");
  B2Q_CHECK_DTYPE(dtype);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == B2Q_C64) return norm2_t<float>(state, n_qubits, out, s);
  return norm2_t<double>(state, n_qubits, out, s);
}

extern "C" int b2q_sv_gather(const void* state, int dtype, int n_qubits, const uint64_t* indices,
                             uint64_t count, double* out_c128, void* stream) {
  B2Q_REQUIRE(state != nullptr && (count == 0 || (indices && out_c128)), "null argument");
  B2Q_CHECK_DTYPE(dtype);
  if (count == 0) return B2Q_OK;
  const uint64_t total = 1ull << n_qubits;
  for (uint64_t j = 0; j < count; ++j)
    B2Q_REQUIRE(indices[j] < total, "index %llu out of range", (unsigned long long)indices[j]);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // 16-byte aligned results first, then the indices
  double2* dout = reinterpret_cast<double2*>(workspace((sizeof(uint64_t) + sizeof(double2)) * count));
  if (dout == nullptr) return B2Q_ERR_CUDA;
  uint64_t* didx = reinterpret_cast<uint64_t*>(dout + count);
  B2Q_CUDA_CHECK(
      cudaMemcpyAsync(didx, indices, sizeof(uint64_t) * count, cudaMemcpyHostToDevice, s));
  const unsigned blocks = (unsigned)((count + 255) / 256);
  if (dtype == B2Q_C64)
    sv_gather_kernel<float><<<blocks, 256, 0, s>>>(reinterpret_cast<const float2*>(state), didx,
                                                   count, dout);
  else
    sv_gather_kernel<double><<<blocks, 256, 0, s>>>(reinterpret_cast<const double2*>(state), didx,
                                                    count, dout);
  B2Q_LAUNCH_CHECK("sv_gather_kernel");
  B2Q_CUDA_CHECK(
      cudaMemcpyAsync(out_c128, dout, sizeof(double2) * count, cudaMemcpyDeviceToHost, s));
  B2Q_CUDA_CHECK(cudaStreamSynchronize(s));
  return B2Q_OK;
}

extern "C" int b2q_sv_marginal_probs(const void* state, int dtype, int n_qubits, const int* bits,
                                     int m, double* probs_dev, double* probs_host, void* stream) {
  B2Q_REQUIRE(state != nullptr && bits != nullptr && probs_dev != nullptr, "null argument");
  B2Q_CHECK_DTYPE(dtype);
  B2Q_REQUIRE(m >= 1 && m <= 24 && m <= n_qubits, "m=%d out of range", m);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  MarginalParams p;
  p.n = n_qubits;
  p.zb = dtype == B2Q_C64 ? 6 : 5;
  p.m = m;
  uint64_t seen = 0;
  for (int q = 0; q < m; ++q) {
    B2Q_REQUIRE(bits[q] >= 0 && bits[q] < n_qubits, "bit out of range");
    B2Q_REQUIRE(!((seen >> bits[q]) & 1ull), "duplicate bit %d", bits[q]);
    seen |= 1ull << bits[q];
    p.bits[q] = bits[q];
  }
  const size_t dim = (size_t)1 << m;
  B2Q_CUDA_CHECK(cudaMemsetAsync(probs_dev, 0, sizeof(double) * dim, s));
  if (n_qubits < p.zb + 1) {
    const uint64_t total = 1ull << n_qubits;
    p.n_meas_high = 0;
    p.log2_iters = 0;
    p.num_warps = 0;
    if (dtype == B2Q_C64)
      sv_marginal_small_kernel<float><<<(unsigned)((total + 127) / 128), 128, 0, s>>>(
          reinterpret_cast<const float2*>(state), p, probs_dev);
    else
      sv_marginal_small_kernel<double><<<(unsigned)((total + 127) / 128), 128, 0, s>>>(
          reinterpret_cast<const double2*>(state), p, probs_dev);
    B2Q_LAUNCH_CHECK("sv_marginal_small_kernel");
  } else {
    std::vector<int> mh;
    for (int b = p.zb; b < n_qubits; ++b)
      if ((seen >> b) & 1ull) mh.push_back(b - p.zb);
    p.n_meas_high = (int)mh.size();
    for (int i = 0; i < p.n_meas_high; ++i) p.meas_high_pos[i] = mh[i];
    const int n_high = n_qubits - p.zb;
    const int n_unmeas = n_high - p.n_meas_high;
    // aim for >= 2^15 warp tasks, <= 2^10 iterations each
    int up_bits = std::max(0, 15 - p.n_meas_high);
    up_bits = std::min(up_bits, n_unmeas);
    up_bits = std::max(up_bits, n_unmeas - 10);
    p.log2_iters = n_unmeas - up_bits;
    p.num_warps = 1ull << (p.n_meas_high + up_bits);
    const uint64_t blocks = (p.num_warps + 7) / 8;
    B2Q_REQUIRE(blocks <= 0x7fffffffull, "grid too large");
    if (dtype == B2Q_C64)
      sv_marginal_kernel<float><<<(unsigned)blocks, 256, 0, s>>>(
          reinterpret_cast<const float2*>(state), p, probs_dev);
    else
      sv_marginal_kernel<double><<<(unsigned)blocks, 256, 0, s>>>(
          reinterpret_cast<const double2*>(state), p, probs_dev);
    B2Q_LAUNCH_CHECK("sv_marginal_kernel");
  }
  if (probs_host != nullptr) {
    B2Q_CUDA_CHECK(
        cudaMemcpyAsync(probs_host, probs_dev, sizeof(double) * dim, cudaMemcpyDeviceToHost, s));
    B2Q_CUDA_CHECK(cudaStreamSynchronize(s));
  }
  return B2Q_OK;
}

static void sample_geometry(int n, int* chunk_log2, int* cpb_log2, uint64_t* nchunks,
                            uint64_t* nblocks) {
  *chunk_log2 = std::min(n, kChunkLog2);
  *nchunks = 1ull << (n - *chunk_log2);
  *cpb_log2 = std::min(n - *chunk_log2, kCpbLog2);
  *nblocks = *nchunks >> *cpb_log2;
}

extern "C" uint64_t b2q_sv_sample_workspace_bytes(int n_qubits, uint64_t reps) {
  (void)reps;
  int cl, bl;
  uint64_t nchunks, nblocks;
  sample_geometry(n_qubits, &cl, &bl, &nchunks, &nblocks);
  return sizeof(double) * (nchunks + 2 * nblocks) + 256;
}

extern "C" int b2q_sv_sample(const void* state, int dtype, int n_qubits,
                             const double* uniforms_dev, uint64_t reps,
                             uint64_t* out_indices_dev, void* workspace,
                             uint64_t workspace_bytes, void* stream) {
  B2Q_REQUIRE(state != nullptr && workspace != nullptr, "null argument");
  B2Q_CHECK_DTYPE(dtype);
  B2Q_REQUIRE(workspace_bytes >= b2q_sv_sample_workspace_bytes(n_qubits, reps),
              "workspace too small");
  if (reps == 0) return B2Q_OK;
  B2Q_REQUIRE(uniforms_dev != nullptr && out_indices_dev != nullptr, "null argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int cl, bl;
  uint64_t nchunks, nblocks;
  sample_geometry(n_qubits, &cl, &bl, &nchunks, &nblocks);
  const uint64_t total = 1ull << n_qubits;
  double* chunk_sums = reinterpret_cast<double*>(workspace);
  double* block_sums = chunk_sums + nchunks;
  double* block_cum = block_sums + nblocks;
  const unsigned g1 = stride_grid(nchunks * 32, 256);
  if (dtype == B2Q_C64)
    sv_chunk_sums_kernel<float><<<g1, 256, 0, s>>>(reinterpret_cast<const float2*>(state), total,
                                                   cl, chunk_sums);
  else
    sv_chunk_sums_kernel<double><<<g1, 256, 0, s>>>(reinterpret_cast<const double2*>(state),
                                                    total, cl, chunk_sums);
  B2Q_LAUNCH_CHECK("sv_chunk_sums_kernel");
  block_sums_kernel<<<stride_grid(nblocks * 32, 256), 256, 0, s>>>(chunk_sums, nblocks, bl,
                                                                   block_sums);
  B2Q_LAUNCH_CHECK("block_sums_kernel");
  scan_inclusive_kernel<<<1, 1024, 0, s>>>(block_sums, block_cum, nblocks);
  B2Q_LAUNCH_CHECK("scan_inclusive_kernel");
  const unsigned g2 = stride_grid(reps * 32, 256);
  if (dtype == B2Q_C64)
    sv_sample_resolve_kernel<float><<<g2, 256, 0, s>>>(reinterpret_cast<const float2*>(state),
                                                       total, cl, bl, chunk_sums, block_cum,
                                                       nblocks, uniforms_dev, reps,
                                                       out_indices_dev);
  else
    sv_sample_resolve_kernel<double><<<g2, 256, 0, s>>>(reinterpret_cast<const double2*>(state),
                                                        total, cl, bl, chunk_sums, block_cum,
                                                        nblocks, uniforms_dev, reps,
                                                        out_indices_dev);
  B2Q_LAUNCH_CHECK("sv_sample_resolve_kernel");
  return B2Q_OK;
}

extern "C" int b2q_cdf_sample(const double* probs_dev, uint64_t count, const double* uniforms_dev,
                              uint64_t reps, uint64_t* out_indices_dev, void* stream) {
  B2Q_REQUIRE(probs_dev != nullptr && count >= 1, "bad distribution");
  if (reps == 0) return B2Q_OK;
  B2Q_REQUIRE(uniforms_dev != nullptr && out_indices_dev != nullptr, "null argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  double* cum = reinterpret_cast<double*>(workspace(sizeof(double) * count));
  if (cum == nullptr) return B2Q_ERR_CUDA;
  scan_inclusive_kernel<<<1, 1024, 0, s>>>(probs_dev, cum, count);
  B2Q_LAUNCH_CHECK("scan_inclusive_kernel");
  cdf_search_kernel<<<(unsigned)((reps + 255) / 256), 256, 0, s>>>(cum, count, uniforms_dev, reps,
                                                                   out_indices_dev);
  B2Q_LAUNCH_CHECK("cdf_search_kernel");
  return B2Q_OK;
}

extern "C" int b2q_unpack_bits(const uint64_t* indices_dev, uint64_t reps, const int* bits, int m,
                               uint8_t* out_dev, void* stream) {
  if (reps == 0 || m == 0) return B2Q_OK;
  B2Q_REQUIRE(indices_dev != nullptr && bits != nullptr && out_dev != nullptr, "null argument");
  B2Q_REQUIRE(m <= 64, "m too large");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  UnpackParams p;
  p.m = m;
  for (int q = 0; q < m; ++q) {
    B2Q_REQUIRE(bits[q] >= 0 && bits[q] < 64, "bit out of range");
    p.bits[q] = bits[q];
  }
  const uint64_t total_out = reps * (uint64_t)m;
  const uint64_t blocks = (total_out + 255) / 256;
  B2Q_REQUIRE(blocks <= 0x7fffffffull, "grid too large");
  unpack_bits_kernel<<<(unsigned)blocks, 256, 0, s>>>(indices_dev, total_out, p, out_dev);
  B2Q_LAUNCH_CHECK("unpack_bits_kernel");
  return B2Q_OK;
}

static int collapse_common(void* state, int dtype, uint64_t total, uint64_t mask, uint64_t want,
                           double scale, cudaStream_t s) {
  return scale_mask_common(state, dtype, total, mask, want, scale, 0.0, s);
}

extern "C" int b2q_sv_collapse(void* state, int dtype, int n_qubits, const int* bits,
                               const int* values, int m, double prob, void* stream) {
  B2Q_REQUIRE(state != nullptr && (m == 0 || (bits && values)), "null argument");
  B2Q_CHECK_DTYPE(dtype);
  B2Q_REQUIRE(prob > 0.0, "collapse onto an outcome of probability %g", prob);
  uint64_t mask = 0, want = 0;
  for (int q = 0; q < m; ++q) {
    B2Q_REQUIRE(bits[q] >= 0 && bits[q] < n_qubits, "bit out of range");
    mask |= 1ull << bits[q];
    if (values[q]) want |= 1ull << bits[q];
  }
  // The reference divides by np.sqrt(probs[result]) computed in the state's
  // real dtype (sim/state_vector.py:318).
  const double scale = dtype == B2Q_C64 ? 1.0 / (double)sqrtf((float)prob) : 1.0 / sqrt(prob);
  return collapse_common(state, dtype, 1ull << n_qubits, mask, want, scale,
                         reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int b2q_dm_collapse(void* rho, int dtype, int n_qubits, const int* bits,
                               const int* values, int m, double prob, void* stream) {
  B2Q_REQUIRE(rho != nullptr && (m == 0 || (bits && values)), "null argument");
  B2Q_CHECK_DTYPE(dtype);
  B2Q_REQUIRE(prob > 0.0, "collapse onto an outcome of probability %g", prob);
  B2Q_REQUIRE(2 * n_qubits <= 40, "density matrix too large");
  uint64_t mask = 0, want = 0;
  for (int q = 0; q < m; ++q) {
    B2Q_REQUIRE(bits[q] >= 0 && bits[q] < n_qubits, "bit out of range");
    mask |= (1ull << bits[q]) | (1ull << (bits[q] + n_qubits));
    if (values[q]) want |= (1ull << bits[q]) | (1ull << (bits[q] + n_qubits));
  }
  return collapse_common(rho, dtype, 1ull << (2 * n_qubits), mask, want, 1.0 / prob,
                         reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int b2q_sv_pauli_expectation(const void* state, int dtype, int n_qubits,
                                        uint64_t x_mask, uint64_t z_mask, double* out_re_im,
                                        void* stream) {
  B2Q_REQUIRE(state != nullptr && out_re_im != nullptr, "null argument");
  B2Q_CHECK_DTYPE(dtype);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const uint64_t total = 1ull << n_qubits;
  B2Q_REQUIRE(x_mask < total && z_mask < total, "mask out of range");
  const unsigned blocks = stride_grid(total, 256);
  double* partial = reinterpret_cast<double*>(workspace(sizeof(double) * 2 * (blocks + 1)));
  if (partial == nullptr) return B2Q_ERR_CUDA;
  if (dtype == B2Q_C64)
    sv_pauli_partial_kernel<float><<<blocks, 256, 0, s>>>(reinterpret_cast<const float2*>(state),
                                                          total, x_mask, z_mask, partial);
  else
    sv_pauli_partial_kernel<double><<<blocks, 256, 0, s>>>(
        reinterpret_cast<const double2*>(state), total, x_mask, z_mask, partial);
  B2Q_LAUNCH_CHECK("sv_pauli_partial_kernel");
  final_sum_kernel<<<1, 256, 0, s>>>(partial, blocks, 2, partial + 2 * blocks);
  B2Q_LAUNCH_CHECK("final_sum_kernel");
  double h[2];
  B2Q_CUDA_CHECK(
      cudaMemcpyAsync(h, partial + 2 * blocks, sizeof(double) * 2, cudaMemcpyDeviceToHost, s));
  B2Q_CUDA_CHECK(cudaStreamSynchronize(s));
  // P|i> = i^{nY} (-1)^{popcount(i & z)} |i ^ x>, nY = popcount(x & z)
  const int ny = __builtin_popcountll(x_mask & z_mask) & 3;
  double re = h[0], im = h[1];
  for (int t = 0; t < ny; ++t) {  // multiply by i
    const double nr = -im, ni = re;
    re = nr;
    im = ni;
  }
  out_re_im[0] = re;
  out_re_im[1] = im;
  return B2Q_OK;
}

extern "C" int b2q_sv_pauli_expectation_multi(const void* state, int dtype, int n_qubits,
                                              uint64_t x_mask, const uint64_t* z_masks, int count,
                                              double* out_re_im, void* stream) {
  B2Q_REQUIRE(state != nullptr && z_masks != nullptr && out_re_im != nullptr, "null argument");
  B2Q_CHECK_DTYPE(dtype);
  B2Q_REQUIRE(count >= 1, "no Pauli strings");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const uint64_t total = 1ull << n_qubits;
  B2Q_REQUIRE(x_mask < total, "mask out of range");
  // CIRQ_B200_PAULI_RUNS=0 keeps the amplitude-per-thread kernel (A/B measurements)
  static const bool runs_enabled = [] {
    const char* e = std::getenv("CIRQ_B200_PAULI_RUNS");
    return e == nullptr || e[0] != '0';
  }();
  const PauliRunPlan plan = pauli_run_plan(n_qubits, runs_enabled);
  const bool by_runs = plan.by_runs;
  const uint64_t vt = plan.vt;
  const int k_count = plan.k_count;
  const unsigned blocks = by_runs ? plan.blocks : stride_grid(total, 256);
  for (int t0 = 0; t0 < count; t0 += kPauliMaxTerms) {
    PauliMultiParams p;
    p.count = std::min(kPauliMaxTerms, count - t0);
    for (int t = 0; t < kPauliMaxTerms; ++t) {
      p.zmask[t] = t < p.count ? z_masks[t0 + t] : 0;
      B2Q_REQUIRE(p.zmask[t] < total, "mask out of range");
    }
    const int width = 2 * p.count;
    double* partial = reinterpret_cast<double*>(workspace(sizeof(double) * width * ((size_t)blocks + 1)));
    if (partial == nullptr) return B2Q_ERR_CUDA;
    if (by_runs && dtype == B2Q_C64 && x_mask == 0)
      sv_pauli_multi_run_kernel<float, true><<<blocks, 256, 0, s>>>(
          reinterpret_cast<const float2*>(state), vt, k_count, x_mask, p, partial);
    else if (by_runs && dtype == B2Q_C64)
      sv_pauli_multi_run_kernel<float, false><<<blocks, 256, 0, s>>>(
          reinterpret_cast<const float2*>(state), vt, k_count, x_mask, p, partial);
    else if (by_runs && x_mask == 0)
      sv_pauli_multi_run_kernel<double, true><<<blocks, 256, 0, s>>>(
          reinterpret_cast<const double2*>(state), vt, k_count, x_mask, p, partial);
    else if (by_runs)
      sv_pauli_multi_run_kernel<double, false><<<blocks, 256, 0, s>>>(
          reinterpret_cast<const double2*>(state), vt, k_count, x_mask, p, partial);
    else if (dtype == B2Q_C64 && x_mask == 0)
      sv_pauli_multi_partial_kernel<float, true><<<blocks, 256, 0, s>>>(
          reinterpret_cast<const float2*>(state), total, x_mask, p, partial);
    else if (dtype == B2Q_C64)
      sv_pauli_multi_partial_kernel<float, false><<<blocks, 256, 0, s>>>(
          reinterpret_cast<const float2*>(state), total, x_mask, p, partial);
    else if (x_mask == 0)
      sv_pauli_multi_partial_kernel<double, true><<<blocks, 256, 0, s>>>(
          reinterpret_cast<const double2*>(state), total, x_mask, p, partial);
    else
      sv_pauli_multi_partial_kernel<double, false><<<blocks, 256, 0, s>>>(
          reinterpret_cast<const double2*>(state), total, x_mask, p, partial);
    B2Q_LAUNCH_CHECK("sv_pauli_multi_partial_kernel");
    column_sum_kernel<<<1, 256, 0, s>>>(partial, blocks, width, partial + (size_t)blocks * width);
    B2Q_LAUNCH_CHECK("column_sum_kernel");
    double h[2 * kPauliMaxTerms];
    B2Q_CUDA_CHECK(cudaMemcpyAsync(h, partial + (size_t)blocks * width, sizeof(double) * width,
                                   cudaMemcpyDeviceToHost, s));
    B2Q_CUDA_CHECK(cudaStreamSynchronize(s));
    for (int t = 0; t < p.count; ++t) {
      // P|i> = i^{nY} (-1)^{popcount(i & z)} |i ^ x>, nY = popcount(x & z)
      const int ny = __builtin_popcountll(x_mask & p.zmask[t]) & 3;
      double re = h[2 * t], im = h[2 * t + 1];
      for (int k = 0; k < ny; ++k) {  // multiply by i
        const double nr = -im, ni = re;
        re = nr;
        im = ni;
      }
      out_re_im[2 * (t0 + t)] = re;
      out_re_im[2 * (t0 + t) + 1] = im;
    }
  }
  return B2Q_OK;
}

extern "C" int b2q_sv_reduced_density_matrix(const void* state, int dtype, int n_qubits,
                                             const int* bits, int m, double* out_c128,
                                             void* stream) {
  B2Q_REQUIRE(state != nullptr && bits != nullptr && out_c128 != nullptr, "null argument");
  B2Q_CHECK_DTYPE(dtype);
  B2Q_REQUIRE(m >= 1 && m <= 5 && m <= n_qubits, "1..5 kept qubits, got %d", m);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  RdmParams p;
  int sorted[6];
  for (int q = 0; q < m; ++q) {
    B2Q_REQUIRE(bits[q] >= 0 && bits[q] < n_qubits, "bit %d out of range", bits[q]);
    sorted[q] = bits[q];
  }
  std::sort(sorted, sorted + m);
  for (int q = 0; q < 6; ++q) p.pos[q] = q < m ? sorted[q] : 0;
  for (int q = 1; q < m; ++q) B2Q_REQUIRE(sorted[q] != sorted[q - 1], "duplicate bit %d", sorted[q]);
  const int d = 1 << m;
  // CIRQ_B200_RDM_GRAM=0 keeps the row-tile kernel at every size (A/B measurements)
  static const bool gram_enabled = [] {
    const char* e = std::getenv("CIRQ_B200_RDM_GRAM");
    return e == nullptr || e[0] != '0';
  }();
  int rc;
  double* out_dev = nullptr;
  if (gram_enabled && m >= 3 && n_qubits >= kRdmTileBits) {
    const RdmGramParams gp = rdm_gram_plan(sorted, m);
    const uint64_t tiles = 1ull << (n_qubits - kRdmTileBits);
    const bool f32 = dtype == B2Q_C64;
    const int ctas = m == 3   ? (f32 ? rdm_gram_grid<float, 3>(tiles) : rdm_gram_grid<double, 3>(tiles))
                     : m == 4 ? (f32 ? rdm_gram_grid<float, 4>(tiles) : rdm_gram_grid<double, 4>(tiles))
                              : (f32 ? rdm_gram_grid<float, 5>(tiles) : rdm_gram_grid<double, 5>(tiles));
    out_dev = reinterpret_cast<double*>(workspace(sizeof(double) * 2 * d * d * (1 + (size_t)ctas)));
    if (out_dev == nullptr) return B2Q_ERR_CUDA;
    double* partial = out_dev + 2 * d * d;
    rc = m == 3   ? (f32 ? reduced_dm_gram_launch<float, 3>(state, n_qubits, gp, ctas, partial, out_dev, s)
                         : reduced_dm_gram_launch<double, 3>(state, n_qubits, gp, ctas, partial, out_dev, s))
         : m == 4 ? (f32 ? reduced_dm_gram_launch<float, 4>(state, n_qubits, gp, ctas, partial, out_dev, s)
                         : reduced_dm_gram_launch<double, 4>(state, n_qubits, gp, ctas, partial, out_dev, s))
                  : (f32 ? reduced_dm_gram_launch<float, 5>(state, n_qubits, gp, ctas, partial, out_dev, s)
                         : reduced_dm_gram_launch<double, 5>(state, n_qubits, gp, ctas, partial, out_dev, s));
  } else {
    out_dev = reinterpret_cast<double*>(workspace(sizeof(double) * 2 * d * d));
    if (out_dev == nullptr) return B2Q_ERR_CUDA;
    B2Q_CUDA_CHECK(cudaMemsetAsync(out_dev, 0, sizeof(double) * 2 * d * d, s));
    rc = dtype == B2Q_C64 ? reduced_dm_t<float>(state, n_qubits, m, p, out_dev, s)
                          : reduced_dm_t<double>(state, n_qubits, m, p, out_dev, s);
  }
  if (rc != B2Q_OK) return rc;
  std::vector<double> h((size_t)2 * d * d);
  B2Q_CUDA_CHECK(cudaMemcpyAsync(h.data(), out_dev, sizeof(double) * 2 * d * d,
                                 cudaMemcpyDeviceToHost, s));
  B2Q_CUDA_CHECK(cudaStreamSynchronize(s));
  // kernel index: sorted positions, LSB first  ->  caller's: bits[0] = MSB
  auto to_kernel = [&](int idx) {
    int k = 0;
    for (int q = 0; q < m; ++q)
      if ((idx >> (m - 1 - q)) & 1)
        for (int i = 0; i < m; ++i)
          if (sorted[i] == bits[q]) k |= 1 << i;
    return k;
  };
  for (int a = 0; a < d; ++a)
    for (int b = 0; b < d; ++b) {
      const size_t src = 2 * ((size_t)to_kernel(a) * d + to_kernel(b));
      out_c128[2 * (a * d + b)] = h[src];
      out_c128[2 * (a * d + b) + 1] = h[src + 1];
    }
  return B2Q_OK;
}

extern "C" int b2q_dm_pauli_expectation(const void* rho, int dtype, int n_qubits, uint64_t x_mask,
                                        uint64_t z_mask, double* out_re_im, void* stream) {
  B2Q_REQUIRE(rho != nullptr && out_re_im != nullptr, "null argument");
  B2Q_CHECK_DTYPE(dtype);
  B2Q_REQUIRE(n_qubits >= 1 && 2 * n_qubits <= 40, "density matrix too large");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const uint64_t dim = 1ull << n_qubits;
  B2Q_REQUIRE(x_mask < dim && z_mask < dim, "mask out of range");
  const unsigned blocks = stride_grid(dim, 256);
  double* partial = reinterpret_cast<double*>(workspace(sizeof(double) * 2 * (blocks + 1)));
  if (partial == nullptr) return B2Q_ERR_CUDA;
  if (dtype == B2Q_C64)
    dm_pauli_partial_kernel<float><<<blocks, 256, 0, s>>>(reinterpret_cast<const float2*>(rho),
                                                          n_qubits, x_mask, z_mask, partial);
  else
    dm_pauli_partial_kernel<double><<<blocks, 256, 0, s>>>(reinterpret_cast<const double2*>(rho),
                                                           n_qubits, x_mask, z_mask, partial);
  B2Q_LAUNCH_CHECK("dm_pauli_partial_kernel");
  final_sum_kernel<<<1, 256, 0, s>>>(partial, blocks, 2, partial + 2 * blocks);
  B2Q_LAUNCH_CHECK("final_sum_kernel");
  double h[2];
  B2Q_CUDA_CHECK(
      cudaMemcpyAsync(h, partial + 2 * blocks, sizeof(double) * 2, cudaMemcpyDeviceToHost, s));
  B2Q_CUDA_CHECK(cudaStreamSynchronize(s));
  // <i ^ x| P |i> = i^{nY} (-1)^{popcount(i & z)}, nY = popcount(x & z)
  const int ny = __builtin_popcountll(x_mask & z_mask) & 3;
  double re = h[0], im = h[1];
  for (int t = 0; t < ny; ++t) {  // multiply by i
    const double nr = -im, ni = re;
    re = nr;
    im = ni;
  }
  out_re_im[0] = re;
  out_re_im[1] = im;
  return B2Q_OK;
}

extern "C" int b2q_dm_diagonal(const void* rho, int dtype, int n_qubits, double* probs_dev,
                               void* stream) {
  B2Q_REQUIRE(rho != nullptr && probs_dev != nullptr, "null argument");
  B2Q_CHECK_DTYPE(dtype);
  B2Q_REQUIRE(2 * n_qubits <= 40, "density matrix too large");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const uint64_t dim = 1ull << n_qubits;
  const unsigned blocks = stride_grid(dim, 256);
  if (dtype == B2Q_C64)
    dm_diag_kernel<float><<<blocks, 256, 0, s>>>(reinterpret_cast<const float2*>(rho), n_qubits,
                                                 probs_dev);
  else
    dm_diag_kernel<double><<<blocks, 256, 0, s>>>(reinterpret_cast<const double2*>(rho), n_qubits,
                                                  probs_dev);
  B2Q_LAUNCH_CHECK("dm_diag_kernel");
  return B2Q_OK;
}

extern "C" int b2q_probs_marginal(const double* probs_dev, int n_qubits, const int* bits, int m,
                                  double* out_dev, void* stream) {
  B2Q_REQUIRE(probs_dev != nullptr && bits != nullptr && out_dev != nullptr, "null argument");
  B2Q_REQUIRE(n_qubits >= 0 && n_qubits <= 30, "n_qubits out of range");
  B2Q_REQUIRE(m >= 1 && m <= 24 && m <= n_qubits, "m=%d out of range", m);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  MarginalParams p;
  p.n = n_qubits;
  p.zb = 0;
  p.m = m;
  p.n_meas_high = 0;
  p.log2_iters = 0;
  p.num_warps = 0;
  uint64_t seen = 0;
  for (int q = 0; q < m; ++q) {
    B2Q_REQUIRE(bits[q] >= 0 && bits[q] < n_qubits, "bit out of range");
    B2Q_REQUIRE(!((seen >> bits[q]) & 1ull), "duplicate bit %d", bits[q]);
    seen |= 1ull << bits[q];
    p.bits[q] = bits[q];
  }
  B2Q_CUDA_CHECK(cudaMemsetAsync(out_dev, 0, sizeof(double) << m, s));
  const uint64_t total = 1ull << n_qubits;
  probs_marginal_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(probs_dev, p, out_dev);
  B2Q_LAUNCH_CHECK("probs_marginal_kernel");
  return B2Q_OK;
}

extern "C" int b2q_dm_trace(const void* rho, int dtype, int n_qubits, double* out, void* stream) {
  B2Q_REQUIRE(rho != nullptr && out != nullptr, "null argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const uint64_t dim = 1ull << n_qubits;
  double* probs = reinterpret_cast<double*>(workspace(sizeof(double) * (dim + 1)));
  if (probs == nullptr) return B2Q_ERR_CUDA;
  int rc = b2q_dm_diagonal(rho, dtype, n_qubits, probs, stream);
  if (rc != B2Q_OK) return rc;
  final_sum_kernel<<<1, 256, 0, s>>>(probs, dim, 1, probs + dim);
  B2Q_LAUNCH_CHECK("final_sum_kernel");
  B2Q_CUDA_CHECK(cudaMemcpyAsync(out, probs + dim, sizeof(double), cudaMemcpyDeviceToHost, s));
  B2Q_CUDA_CHECK(cudaStreamSynchronize(s));
  return B2Q_OK;
}

static int pack_common(void* shard, int dtype, int n_local, const int* local_bits, int g,
                       void* packed, int dir, void* stream) {
  B2Q_REQUIRE(shard != nullptr && packed != nullptr && local_bits != nullptr, "null argument");
  B2Q_CHECK_DTYPE(dtype);
  B2Q_REQUIRE(g >= 1 && g <= 8 && g <= n_local, "g=%d out of range", g);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  PackParams p;
  p.n_local = n_local;
  p.g = g;
  // Segment index: local_bits[0] = MSB.
  std::vector<std::pair<int, int>> pr;
  for (int i = 0; i < g; ++i) {
    B2Q_REQUIRE(local_bits[i] >= 0 && local_bits[i] < n_local, "local bit out of range");
    pr.push_back({local_bits[i], g - 1 - i});
  }
  std::sort(pr.begin(), pr.end());
  for (int i = 0; i < g; ++i) {
    B2Q_REQUIRE(i == 0 || pr[i].first != pr[i - 1].first, "duplicate local bit");
    p.pos[i] = pr[i].first;
    p.rank_of[i] = pr[i].second;
  }
  const uint64_t total = 1ull << n_local;
  const unsigned blocks = stride_grid(total, 256);
  if (dtype == B2Q_C64) {
    if (dir == 0)
      dist_pack_kernel<float, 0><<<blocks, 256, 0, s>>>(reinterpret_cast<float2*>(shard),
                                                        reinterpret_cast<float2*>(packed), p);
    else
      dist_pack_kernel<float, 1><<<blocks, 256, 0, s>>>(reinterpret_cast<float2*>(shard),
                                                        reinterpret_cast<float2*>(packed), p);
  } else {
    if (dir == 0)
      dist_pack_kernel<double, 0><<<blocks, 256, 0, s>>>(reinterpret_cast<double2*>(shard),
                                                         reinterpret_cast<double2*>(packed), p);
    else
      dist_pack_kernel<double, 1><<<blocks, 256, 0, s>>>(reinterpret_cast<double2*>(shard),
                                                         reinterpret_cast<double2*>(packed), p);
  }
  B2Q_LAUNCH_CHECK("dist_pack_kernel");
  return B2Q_OK;
}

extern "C" int b2q_dist_pack(const void* shard, int dtype, int n_local, const int* local_bits,
                             int g, void* packed, void* stream) {
  return pack_common(const_cast<void*>(shard), dtype, n_local, local_bits, g, packed, 0, stream);
}

extern "C" int b2q_dist_unpack(void* shard, int dtype, int n_local, const int* local_bits, int g,
                               const void* packed, void* stream) {
  return pack_common(shard, dtype, n_local, local_bits, g, const_cast<void*>(packed), 1, stream);
}

// Host-only (tests/test_plan_host.py): the Gram kernel's tile plan for kept bits `bits`
// (any order) — out[0..10] tile bits, out[11..15] rank of kept bit q in the tile, then for
// each of the 2^11 tile elements e: out[16 + 2e] = offset from the tile base,
// out[17 + 2e] = slot in the staged tile X[rest][a].
extern "C" int b2q_debug_rdm_plan(int n_qubits, const int* bits, int m, int64_t* out) {
  B2Q_REQUIRE(bits != nullptr && out != nullptr, "null argument");
  B2Q_REQUIRE(m >= 3 && m <= 5 && n_qubits >= kRdmTileBits && n_qubits <= 40, "not a Gram-kernel shape");
  int sorted[5];
  for (int q = 0; q < m; ++q) {
    B2Q_REQUIRE(bits[q] >= 0 && bits[q] < n_qubits, "bit %d out of range", bits[q]);
    sorted[q] = bits[q];
  }
  std::sort(sorted, sorted + m);
  for (int q = 1; q < m; ++q) B2Q_REQUIRE(sorted[q] != sorted[q - 1], "duplicate bit %d", sorted[q]);
  const RdmGramParams gp = rdm_gram_plan(sorted, m);
  for (int j = 0; j < kRdmTileBits; ++j) out[j] = gp.tile_pos[j];
  for (int q = 0; q < 5; ++q) out[kRdmTileBits + q] = gp.kept_rank[q];
  for (int e = 0; e < (1 << kRdmTileBits); ++e) {
    out[16 + 2 * e] = (int64_t)rdm_tile_offset(gp, e);
    out[17 + 2 * e] = rdm_tile_slot(gp, m, e);
  }
  return B2Q_OK;
}

// Host-only: launch shape of b2q_sv_pauli_expectation_multi at this register size —
// out[0] = 1 when the run kernel is used, out[1] = virtual threads, out[2] = runs per
// virtual thread (steps of the sign table), out[3] = CTAs.
extern "C" int b2q_debug_pauli_plan(int n_qubits, int64_t* out) {
  B2Q_REQUIRE(out != nullptr && n_qubits >= 1 && n_qubits <= 40, "bad argument");
  const PauliRunPlan plan = pauli_run_plan(n_qubits, true);
  out[0] = plan.by_runs ? 1 : 0;
  out[1] = (int64_t)plan.vt;
  out[2] = plan.k_count;
  out[3] = plan.blocks;
  return B2Q_OK;
}
