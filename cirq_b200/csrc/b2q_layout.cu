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
#include <cstring>
#include <vector>

namespace b2q {

template <typename real, int VEC>
struct AmpVec;
template <>
struct AmpVec<float, 1> {
  using type = float2;
  static __device__ __forceinline__ float2 at(const float2& v, int) { return v; }
  static __device__ __forceinline__ float2 pack(const float2* e) { return e[0]; }
};
template <>
struct AmpVec<float, 2> {
  using type = float4;
  static __device__ __forceinline__ float2 at(const float4& v, int e) {
    return e == 0 ? make_float2(v.x, v.y) : make_float2(v.z, v.w);
  }
  static __device__ __forceinline__ float4 pack(const float2* e) {
    return make_float4(e[0].x, e[0].y, e[1].x, e[1].y);
  }
};
template <>
struct AmpVec<double, 1> {
  using type = double2;
  static __device__ __forceinline__ double2 at(const double2& v, int) { return v; }
  static __device__ __forceinline__ double2 pack(const double2* e) { return e[0]; }
};
constexpr int kLayoutUnroll = 4;  // independent accesses in flight per thread and operand

// 16-byte accesses of complex64 pairs need aligned operands (sub-states live at any
// 8-byte offset of an arena).
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// out[(i << nb) | j] = a[i] * b[j]; VEC amplitudes of b and out per access (16-byte
// stores: VEC = 2 for complex64 needs nb >= 1 and aligned operands)
template <typename real, int VEC>
__global__ void __launch_bounds__(256)
    sv_kron_kernel(const typename Cplx<real>::type* __restrict__ a,
                   const typename Cplx<real>::type* __restrict__ b, int nb,
                   typename Cplx<real>::type* __restrict__ out, uint64_t total) {
  using C = typename Cplx<real>::type;
  using V = typename AmpVec<real, VEC>::type;
  const uint64_t mask = (1ull << nb) - 1ull;
  const uint64_t nvec = total / VEC, stride = (uint64_t)gridDim.x * blockDim.x;
  V* outv = reinterpret_cast<V*>(out);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += stride * kLayoutUnroll) {
    V y[kLayoutUnroll];
    C x[kLayoutUnroll];
#pragma unroll
    for (int u = 0; u < kLayoutUnroll; ++u) {
      const uint64_t o = (i + u * stride) * VEC;
      if (o < total) {
        x[u] = a[o >> nb];
        y[u] = *reinterpret_cast<const V*>(b + (o & mask));
      }
    }
#pragma unroll
    for (int u = 0; u < kLayoutUnroll; ++u) {
      if ((i + u * stride) * VEC < total) {
        C r[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const C ye = AmpVec<real, VEC>::at(y[u], e);
          r[e] = make_c<real>(x[u].x * ye.x - x[u].y * ye.y, x[u].x * ye.y + x[u].y * ye.x);
        }
        outv[i + u * stride] = AmpVec<real, VEC>::pack(r);
      }
    }
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

// ---- in-place bit permutation, one 64 KB tile per CTA --------------------------
//
// A permutation that only moves index bits inside a set T (|T| <= 13 for complex64,
// 12 for complex128) maps every "tile" — the 2^|T| amplitudes that agree on all
// other bits — onto itself, so it can run IN PLACE: a CTA reads its tile with
// coalesced 16-byte accesses, drops every amplitude at its permuted position in
// shared memory, and writes the tile back linearly.  T always contains the lowest
// index bits, so HBM moves in runs of >= 128 bytes whatever is permuted.  Any
// permutation is a short product of such passes (the host picks them: each pass
// settles up to |T| - 4 more bits).  Replaces np.moveaxis / transpose of
// linalg/transformations.py:743-754 without the second buffer, which is what lets
// states above 30 qubits keep the reference's product-state form
// (sim/simulation_product_state.py:68-81) and puts relabelled SWAPs back.
constexpr int kPermMaxTileBits = 13;

struct PermuteTileParams {
  uint64_t num_tiles;
  int tile_bits;
  int tbits[kPermMaxTileBits];      // ascending index positions of the tile bits, tbits[0] = 0
  int src_local[kPermMaxTileBits];  // output local bit k = input local bit src_local[k]
  int dst_local[kPermMaxTileBits];  // input local bit j goes to output local bit dst_local[j]
  uint32_t xmask[3];                // slot bit d + 1 = parity(position & xmask[d])
};

// Shared-memory slot of tile position o: bits 1-3 are parities chosen on the host
// (`permute_swizzle`) so that the permuted writes of a half warp spread over 8 bank
// pairs (lanes drive position bits dst_local[1..4]; bit 0 stays, for the 16-byte
// linear reads) while the linear reads stay conflict free.
__device__ __forceinline__ uint32_t perm_slot(uint32_t o, const uint32_t* xmask) {
  uint32_t s = o & ~0xeu;
#pragma unroll
  for (int d = 0; d < 3; ++d) s |= (uint32_t)(__popc(o & xmask[d]) & 1) << (d + 1);
  return s;
}

template <typename real>
__global__ void __launch_bounds__(256, 3)
    sv_permute_tile_kernel(typename Cplx<real>::type* __restrict__ state,
                           const __grid_constant__ PermuteTileParams p) {
  using C = typename Cplx<real>::type;
  constexpr int kVec = 16 / sizeof(C);  // amplitudes per 16-byte access
  constexpr int kPerRound = 256 * kVec;  // local indices covered by one round of the CTA
  extern __shared__ __align__(16) unsigned char perm_smem[];
  C* tile = reinterpret_cast<C*>(perm_smem);
  // A local index is (round << log2(kPerRound)) | (thread part): its state offset and its
  // permuted position split the same way (sums / ORs of per-bit contributions), so the
  // per-bit loops run once per thread and once per round, not once per element.
  __shared__ uint64_t round_off[16];
  __shared__ uint32_t round_o[16];
  const uint32_t tile_elems = 1u << p.tile_bits;
  const uint32_t rounds = (tile_elems + kPerRound - 1) / kPerRound;
  const uint32_t l_thread = threadIdx.x * kVec;
  uint64_t off_thread = 0;
  uint32_t o_thread = 0;
  for (int j = 0; j < p.tile_bits; ++j) {
    const uint32_t bit = (l_thread >> j) & 1u;
    off_thread += (uint64_t)bit << p.tbits[j];
    o_thread |= bit << p.dst_local[j];
  }
  if (threadIdx.x < rounds) {
    const uint32_t l = threadIdx.x * kPerRound;
    uint64_t off = 0;
    uint32_t o = 0;
    for (int j = 0; j < p.tile_bits; ++j) {
      const uint32_t bit = (l >> j) & 1u;
      off += (uint64_t)bit << p.tbits[j];
      o |= bit << p.dst_local[j];
    }
    round_off[threadIdx.x] = off;
    round_o[threadIdx.x] = o;
  }
  __syncthreads();
  const bool active = l_thread < tile_elems;
  const uint32_t second = 1u << p.dst_local[0];  // where the vector's second amplitude goes
  constexpr int kBatch = 8;  // global loads in flight per thread
  for (uint64_t t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
    const uint64_t base = insert_zero_bits(t, p.tbits, p.tile_bits) + off_thread;
    if (active) {
      for (uint32_t r0 = 0; r0 < rounds; r0 += kBatch) {
        if constexpr (kVec == 2) {
          float4 v[kBatch];
#pragma unroll
          for (int k = 0; k < kBatch; ++k)
            if (r0 + k < rounds) v[k] = *reinterpret_cast<const float4*>(state + base + round_off[r0 + k]);
#pragma unroll
          for (int k = 0; k < kBatch; ++k)
            if (r0 + k < rounds) {
              const uint32_t o = o_thread | round_o[r0 + k];
              tile[perm_slot(o, p.xmask)] = make_float2(v[k].x, v[k].y);
              tile[perm_slot(o | second, p.xmask)] = make_float2(v[k].z, v[k].w);
            }
        } else {
          C v[kBatch];
#pragma unroll
          for (int k = 0; k < kBatch; ++k)
            if (r0 + k < rounds) v[k] = state[base + round_off[r0 + k]];
#pragma unroll
          for (int k = 0; k < kBatch; ++k)
            if (r0 + k < rounds) tile[perm_slot(o_thread | round_o[r0 + k], p.xmask)] = v[k];
        }
      }
    }
    __syncthreads();
    if (active) {
      for (uint32_t r = 0; r < rounds; ++r) {
        const uint32_t l = l_thread + r * kPerRound;
        if constexpr (kVec == 2) {
          const float4 v = *reinterpret_cast<const float4*>(&tile[perm_slot(l, p.xmask)]);
          *reinterpret_cast<float4*>(state + base + round_off[r]) = v;
        } else {
          state[base + round_off[r]] = tile[perm_slot(l, p.xmask)];
        }
      }
    }
    __syncthreads();
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

// One element of np.allclose: |got - want| > atol + rtol * |want|.  Differences below
// atol (every element of a matching pair of states) are settled without the two
// float64 square roots that bound these kernels before (1.4 TB/s at 28 qubits).
__device__ __forceinline__ bool allclose_violated(double got_re, double got_im, double want_re,
                                                  double want_im, double atol, double atol2,
                                                  double rtol) {
  const double dr = got_re - want_re, di = got_im - want_im;
  const double d2 = dr * dr + di * di;
  if (d2 <= atol2) return false;
  return sqrt(d2) > atol + rtol * sqrt(want_re * want_re + want_im * want_im);
}


// flag[0] = 1 if some |a[i]*b[j] - t[(i<<nb)|j]| > atol + rtol*|t|  (np.allclose);
// VEC amplitudes of t and b per access (VEC = 2 needs nb >= 1 and 16-byte alignment).
template <typename real, int VEC>
__global__ void __launch_bounds__(256)
    sv_kron_mismatch_kernel(const typename Cplx<real>::type* __restrict__ a,
                            const typename Cplx<real>::type* __restrict__ b, int nb,
                            const typename Cplx<real>::type* __restrict__ t, uint64_t total,
                            double atol, double rtol, int* flag) {
  using V = typename AmpVec<real, VEC>::type;
  const uint64_t mask = (1ull << nb) - 1ull;
  const uint64_t nvec = total / VEC, stride = (uint64_t)gridDim.x * blockDim.x;
  const V* tv = reinterpret_cast<const V*>(t);
  const double atol2 = atol * atol;
  bool bad = false;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += stride * kLayoutUnroll) {
    if (*reinterpret_cast<volatile const int*>(flag)) return;  // settled elsewhere
    V z[kLayoutUnroll], y[kLayoutUnroll];
    typename Cplx<real>::type x[kLayoutUnroll];
#pragma unroll
    for (int u = 0; u < kLayoutUnroll; ++u) {
      const uint64_t o = (i + u * stride) * VEC;
      if (o < total) {
        z[u] = tv[i + u * stride];
        y[u] = *reinterpret_cast<const V*>(b + (o & mask));
        x[u] = a[o >> nb];
      }
    }
#pragma unroll
    for (int u = 0; u < kLayoutUnroll; ++u) {
      if ((i + u * stride) * VEC < total) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const auto ye = AmpVec<real, VEC>::at(y[u], e);
          const auto ze = AmpVec<real, VEC>::at(z[u], e);
          if (allclose_violated((double)(x[u].x * ye.x - x[u].y * ye.y),
                                (double)(x[u].x * ye.y + x[u].y * ye.x), (double)ze.x,
                                (double)ze.y, atol, atol2, rtol))
            bad = true;
        }
      }
    }
    if (bad) break;
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
template <typename real, int VEC>
__global__ void __launch_bounds__(256)
    sv_mismatch_kernel(const typename Cplx<real>::type* __restrict__ a,
                       const typename Cplx<real>::type* __restrict__ b, uint64_t total,
                       double atol, double rtol, int* flag) {
  using V = typename AmpVec<real, VEC>::type;
  const uint64_t nvec = total / VEC, stride = (uint64_t)gridDim.x * blockDim.x;
  const V* av = reinterpret_cast<const V*>(a);
  const V* bv = reinterpret_cast<const V*>(b);
  const double atol2 = atol * atol;
  bool bad = false;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += stride * kLayoutUnroll) {
    if (*reinterpret_cast<volatile const int*>(flag)) return;  // settled elsewhere
    V x[kLayoutUnroll], y[kLayoutUnroll];
#pragma unroll
    for (int u = 0; u < kLayoutUnroll; ++u)
      if (i + u * stride < nvec) {
        x[u] = av[i + u * stride];
        y[u] = bv[i + u * stride];
      }
#pragma unroll
    for (int u = 0; u < kLayoutUnroll; ++u)
      if (i + u * stride < nvec) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const auto xe = AmpVec<real, VEC>::at(x[u], e);
          const auto ye = AmpVec<real, VEC>::at(y[u], e);
          if (allclose_violated((double)xe.x, (double)xe.y, (double)ye.x, (double)ye.y, atol,
                                atol2, rtol))
            bad = true;
        }
      }
    if (bad) break;
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
  if (dtype == B2Q_C64 && nb >= 1 && aligned16(b) && aligned16(out))
    sv_kron_kernel<float, 2><<<layout_grid(total / 2 / kLayoutUnroll), 256, 0, s>>>(
        reinterpret_cast<const float2*>(a), reinterpret_cast<const float2*>(b), nb,
        reinterpret_cast<float2*>(out), total);
  else if (dtype == B2Q_C64)
    sv_kron_kernel<float, 1><<<layout_grid(total / kLayoutUnroll), 256, 0, s>>>(
        reinterpret_cast<const float2*>(a), reinterpret_cast<const float2*>(b), nb,
        reinterpret_cast<float2*>(out), total);
  else
    sv_kron_kernel<double, 1><<<layout_grid(total / kLayoutUnroll), 256, 0, s>>>(
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

// Swizzle of a pass: lanes 0-15 of a warp drive input local bits first..first+3
// (first = 1 for complex64, whose lanes move 2 amplitudes; 0 for complex128), i.e. tile
// position bits dst_local[first..first+3].  Each of the slot bits 1-3 must be driven by
// a different one of them: bit d itself when it is among them, else one of the others
// is XORed in.
static void permute_swizzle(PermuteTileParams* p, int first) {
  int lane_bits[4], nl = 0;
  for (int j = first; j < first + 4 && j < p->tile_bits; ++j) lane_bits[nl++] = p->dst_local[j];
  bool used[4] = {false, false, false, false};
  for (int d = 1; d <= 3; ++d) {
    p->xmask[d - 1] = 1u << d;
    for (int i = 0; i < nl; ++i)
      if (lane_bits[i] == d) used[i] = true;
  }
  for (int d = 1; d <= 3; ++d) {
    bool direct = false;
    for (int i = 0; i < nl; ++i) direct |= lane_bits[i] == d;
    if (direct) continue;
    for (int i = 0; i < nl; ++i)
      if (!used[i] && lane_bits[i] >= 4) {
        p->xmask[d - 1] |= 1u << lane_bits[i];
        used[i] = true;
        break;
      }
  }
}

// Host planner of the in-place permutation: `want[k]` = the bit whose content must end
// at position k.  Fills passes (tile bits + local source map) until done.
struct PermutePass {
  int count;
  int tbits[kPermMaxTileBits];
  int src_local[kPermMaxTileBits];
};

static int plan_permute_passes(int n, const int* want, int cap, std::vector<PermutePass>* passes) {
  std::vector<int> cur(n);  // cur[pos] = original bit whose content sits at pos now
  for (int i = 0; i < n; ++i) cur[i] = i;
  const int forced = std::min(n, 4);  // index bits 0-3: runs of >= 128 bytes
  for (int guard = 0; guard < 64; ++guard) {
    bool done = true;
    for (int k = 0; k < n; ++k) done &= cur[k] == want[k];
    if (done) return B2Q_OK;
    std::vector<char> in(n, 0);
    std::vector<int> set;
    auto add = [&](int b) {
      if (!in[b]) {
        in[b] = 1;
        set.push_back(b);
      }
    };
    for (int b = 0; b < forced; ++b) add(b);
    std::vector<int> where(n);  // where[original bit] = current position
    for (int pos = 0; pos < n; ++pos) where[cur[pos]] = pos;
    // follow chains: an unsettled position, then the position holding what it wants, ...
    for (int k = 0; k < n && (int)set.size() < cap; ++k) {
      if (cur[k] == want[k] || in[k]) continue;
      int pos = k;
      while (true) {
        if (!in[pos]) {
          if ((int)set.size() >= cap) break;
          add(pos);
        }
        const int next = where[want[pos]];
        if (in[next]) break;  // the chain closes inside the set
        pos = next;
      }
    }
    // pad with the lowest bits that stay where they are: a full tile keeps all 256
    // threads busy and moves the state in the longest runs the permutation allows
    for (int b = 0; b < n && (int)set.size() < cap; ++b) add(b);
    std::sort(set.begin(), set.end());
    const int T = (int)set.size();
    // sigma on the set: position k takes the content it wants when that content is
    // inside the set; leftovers are matched in order
    std::vector<int> src_pos(n, -1);
    std::vector<char> taken(n, 0);
    for (int k : set) {
      const int q = where[want[k]];
      if (in[q]) {
        src_pos[k] = q;
        taken[q] = 1;
      }
    }
    std::vector<int> free_src;
    for (int q : set)
      if (!taken[q]) free_src.push_back(q);
    size_t fi = 0;
    for (int k : set)
      if (src_pos[k] < 0) src_pos[k] = free_src[fi++];
    PermutePass pass;
    pass.count = T;
    bool moves = false;
    for (int i = 0; i < T; ++i) {
      pass.tbits[i] = set[i];
      const int q = src_pos[set[i]];
      pass.src_local[i] = (int)(std::lower_bound(set.begin(), set.end(), q) - set.begin());
      moves |= q != set[i];
    }
    if (!moves) return set_error(B2Q_ERR_INVALID, "permutation planner made no progress");
    std::vector<int> next_cur(cur);
    for (int k : set) next_cur[k] = cur[src_pos[k]];
    cur.swap(next_cur);
    passes->push_back(pass);
  }
  return set_error(B2Q_ERR_INVALID, "permutation planner did not converge");
}

extern "C" int b2q_sv_permute_bits_inplace(void* state, int dtype, int n_qubits, const int* src_bit,
                                           int* passes_out, void* stream) {
  B2Q_REQUIRE(state != nullptr && src_bit != nullptr, "null argument");
  B2Q_REQUIRE(dtype == B2Q_C64 || dtype == B2Q_C128, "bad dtype %d", dtype);
  B2Q_REQUIRE(n_qubits >= 1 && n_qubits <= 40, "n_qubits out of range");
  uint64_t seen = 0;
  for (int k = 0; k < n_qubits; ++k) {
    B2Q_REQUIRE(src_bit[k] >= 0 && src_bit[k] < n_qubits, "source bit out of range");
    B2Q_REQUIRE(!((seen >> src_bit[k]) & 1ull), "src_bit is not a permutation");
    seen |= 1ull << src_bit[k];
  }
  const int cap = std::min(n_qubits, dtype == B2Q_C64 ? kPermMaxTileBits : kPermMaxTileBits - 1);
  std::vector<PermutePass> passes;
  const int rc = plan_permute_passes(n_qubits, src_bit, cap, &passes);
  if (rc != B2Q_OK) return rc;
  if (passes_out != nullptr) *passes_out = (int)passes.size();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  static bool attr_set = false;
  if (!attr_set) {
    B2Q_CUDA_CHECK(cudaFuncSetAttribute(sv_permute_tile_kernel<float>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    B2Q_CUDA_CHECK(cudaFuncSetAttribute(sv_permute_tile_kernel<double>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    attr_set = true;
  }
  for (const PermutePass& pass : passes) {
    PermuteTileParams p;
    memset(&p, 0, sizeof(p));
    p.tile_bits = pass.count;
    p.num_tiles = 1ull << (n_qubits - pass.count);
    for (int i = 0; i < pass.count; ++i) {
      p.tbits[i] = pass.tbits[i];
      p.src_local[i] = pass.src_local[i];
      p.dst_local[pass.src_local[i]] = i;
    }
    B2Q_REQUIRE(p.tbits[0] == 0, "tile must contain index bit 0");
    permute_swizzle(&p, dtype == B2Q_C64 ? 1 : 0);
    const size_t smem = elem_bytes(dtype) << pass.count;
    const unsigned grid = (unsigned)std::min<uint64_t>(p.num_tiles, 148ull * 3);
    if (dtype == B2Q_C64)
      sv_permute_tile_kernel<float><<<grid, 256, smem, s>>>(reinterpret_cast<float2*>(state), p);
    else
      sv_permute_tile_kernel<double><<<grid, 256, smem, s>>>(reinterpret_cast<double2*>(state), p);
    B2Q_LAUNCH_CHECK("sv_permute_tile_kernel");
  }
  return B2Q_OK;
}

// Host-only: the passes b2q_sv_permute_bits_inplace would run.  out: per pass
// count | tbits[13] | src_local[13] (27 ints), at most max_passes of them;
// returns the number of passes in *passes_out.
extern "C" int b2q_debug_permute_plan(int dtype, int n_qubits, const int* src_bit, int max_passes,
                                      int* out, int* passes_out) {
  B2Q_REQUIRE(src_bit != nullptr && out != nullptr && passes_out != nullptr, "null argument");
  const int cap = std::min(n_qubits, dtype == B2Q_C64 ? kPermMaxTileBits : kPermMaxTileBits - 1);
  std::vector<PermutePass> passes;
  const int rc = plan_permute_passes(n_qubits, src_bit, cap, &passes);
  if (rc != B2Q_OK) return rc;
  *passes_out = (int)passes.size();
  for (int i = 0; i < (int)passes.size() && i < max_passes; ++i) {
    int* o = out + 27 * i;
    o[0] = passes[i].count;
    for (int j = 0; j < kPermMaxTileBits; ++j) {
      o[1 + j] = j < passes[i].count ? passes[i].tbits[j] : -1;
      o[14 + j] = j < passes[i].count ? passes[i].src_local[j] : -1;
    }
  }
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
  if (dtype == B2Q_C64 && nb >= 1 && aligned16(b) && aligned16(t))
    sv_kron_mismatch_kernel<float, 2><<<layout_grid(total / 2 / kLayoutUnroll), 256, 0, s>>>(
        reinterpret_cast<const float2*>(a), reinterpret_cast<const float2*>(b), nb,
        reinterpret_cast<const float2*>(t), total, atol, rtol, flag);
  else if (dtype == B2Q_C64)
    sv_kron_mismatch_kernel<float, 1><<<layout_grid(total / kLayoutUnroll), 256, 0, s>>>(
        reinterpret_cast<const float2*>(a), reinterpret_cast<const float2*>(b), nb,
        reinterpret_cast<const float2*>(t), total, atol, rtol, flag);
  else
    sv_kron_mismatch_kernel<double, 1><<<layout_grid(total / kLayoutUnroll), 256, 0, s>>>(
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
  if (dtype == B2Q_C64 && n_qubits >= 1 && aligned16(a) && aligned16(b))
    sv_mismatch_kernel<float, 2><<<layout_grid(total / 2 / kLayoutUnroll), 256, 0, s>>>(
        reinterpret_cast<const float2*>(a), reinterpret_cast<const float2*>(b), total, atol, rtol,
        flag);
  else if (dtype == B2Q_C64)
    sv_mismatch_kernel<float, 1><<<layout_grid(total / kLayoutUnroll), 256, 0, s>>>(
        reinterpret_cast<const float2*>(a), reinterpret_cast<const float2*>(b), total, atol, rtol,
        flag);
  else
    sv_mismatch_kernel<double, 1><<<layout_grid(total / kLayoutUnroll), 256, 0, s>>>(
        reinterpret_cast<const double2*>(a), reinterpret_cast<const double2*>(b), total, atol,
        rtol, flag);
  B2Q_LAUNCH_CHECK("sv_mismatch_kernel");
  int h = 0;
  B2Q_CUDA_CHECK(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, s));
  B2Q_CUDA_CHECK(cudaStreamSynchronize(s));
  *ok_out = h ? 0 : 1;
  return B2Q_OK;
}
