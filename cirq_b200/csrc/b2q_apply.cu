// Gate application on a dense state vector: psi <- (M on target bits) psi.
//
// Replaces (reference, cirq-core/cirq/): linalg/transformations.py:105-172
// (targeted_left_multiply, one np.einsum over the whole state), :310-380
// (apply_matrix_to_slices) and the per-gate slice tricks of ops/*.py, as
// reached from protocols/apply_unitary_protocol.py:440-466.
//
// Design: every gate is ONE in-place streaming pass.
// A warp owns a "zone" of 512 contiguous bytes (the low 6 bits of the index for
// complex64, 5 for complex128) times 2^h strided copies of it, one per
// combination of the h register-resident high bits.  Each lane moves 16 bytes
// per access, so every global load/store instruction of a warp covers 512
// contiguous bytes regardless of which qubits the gate acts on:
//   * target bits above the zone   -> different registers of the same thread
//     (coalesced strided copies),
//   * target bit 0 (complex64)     -> the two halves of the 16-byte vector,
//   * target bits inside the zone  -> exchanged with a spare register bit by a
//     warp-shuffle butterfly before the multiply and back after it.
// After the exchange every thread holds complete 2^k-amplitude groups in
// registers; the matrix sits in the kernel parameter (constant) bank so the
// FMAs take it as a direct operand: no shared memory, no barriers.
#include "b2q_common.cuh"

#include <algorithm>
#include <cstring>
#include <vector>

namespace b2q {

constexpr int kMaxRegBits = 6;
constexpr int kMaxIns = 6;
constexpr int kFastThreads = 256;

// Matrix entries in the kernel parameter bank.  complex128: (re, im).
// complex64: (re, re, -im, im), the two packed operands of the Blackwell FFMA2
// (fma.rn.f32x2) form of a complex multiply-accumulate:
//   acc(re,im) += (re,re) * x(re,im);  acc(re,im) += (-im,im) * x(im,re)
// — two instructions per complex MAC instead of four, the (im,re) swap being an
// operand modifier in SASS (R.F32x2.LO_HI).
template <typename real>
struct MatEntry {
  static constexpr int kReals = sizeof(real) == 4 ? 4 : 2;
};

template <typename real, typename C>
__device__ __forceinline__ void cmac_entry(C& acc, const real* m, const C& x);
template <>
__device__ __forceinline__ void cmac_entry<float, float2>(float2& acc, const float* m,
                                                          const float2& x) {
  acc = __ffma2_rn(make_float2(m[0], m[1]), x, acc);
  acc = __ffma2_rn(make_float2(m[2], m[3]), make_float2(x.y, x.x), acc);
}
template <>
__device__ __forceinline__ void cmac_entry<double, double2>(double2& acc, const double* m,
                                                            const double2& x) {
  cmac<double>(acc, m[0], m[1], x);
}

template <typename real, int K>
struct FastParams {
  typename Cplx<real>::type* state;
  uint64_t num_items;               // warp-items (one 512-byte zone x 2^h copies each)
  int n_ins;                        // number of register-resident high bits
  int ins_pos[kMaxIns];             // their positions, ascending
  long long reg_off[kMaxRegBits];   // element offset contributed by each register bit
  int swap_lane[kMaxRegBits];       // per target slot: lane bit to exchange with, or -1
  real mat[MatEntry<real>::kReals << (2 * K)];  // row-major entries (see MatEntry)
};

template <typename real, int K>
struct SmallParams {
  typename Cplx<real>::type* state;
  uint64_t num_groups;
  int tpos[K];  // ascending
  real mat[MatEntry<real>::kReals << (2 * K)];
};


// 16-byte global accesses.  `volatile` keeps the loads of a thread grouped in
// program order ahead of the arithmetic so all of them are in flight at once.
__device__ __forceinline__ float4 ldg16(const float2* p) {
  float4 v;
  asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ float2 ldg_elem(const float2* p) {
  float2 v;
  asm volatile("ld.global.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_elem(float2* p, const float2& v) {
  asm volatile("st.global.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ double2 ldg16(const double2* p) {
  double2 v;
  asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg16(float2* p, const float4& v) {
  asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void stg16(double2* p, const double2& v) {
  asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

__device__ __forceinline__ double2 ldg_elem(const double2* p) { return ldg16(p); }
__device__ __forceinline__ void stg_elem(double2* p, const double2& v) { stg16(p, v); }

template <typename C>
__device__ __forceinline__ C shfl_xor_c(const C& v, int mask);
template <>
__device__ __forceinline__ float2 shfl_xor_c<float2>(const float2& v, int mask) {
  float2 r;
  r.x = __shfl_xor_sync(0xffffffffu, v.x, mask);
  r.y = __shfl_xor_sync(0xffffffffu, v.y, mask);
  return r;
}
template <>
__device__ __forceinline__ double2 shfl_xor_c<double2>(const double2& v, int mask) {
  double2 r;
  r.x = __shfl_xor_sync(0xffffffffu, v.x, mask);
  r.y = __shfl_xor_sync(0xffffffffu, v.y, mask);
  return r;
}

// acc[rr] = sum_c M[r0+rr][c] * xin(c), rr < RT: RT independent accumulator chains.
template <typename real, typename C, int DIM, int RT, typename XF>
__device__ __forceinline__ void rows_tile(const real* __restrict__ mat, int r0, C (&acc)[RT],
                                          XF xin) {
  constexpr int ME = MatEntry<real>::kReals;
#pragma unroll
  for (int rr = 0; rr < RT; ++rr) acc[rr] = make_c<real>(0, 0);
#pragma unroll
  for (int c = 0; c < DIM; ++c) {
    const C xv = xin(c);
#pragma unroll
    for (int rr = 0; rr < RT; ++rr) cmac_entry<real, C>(acc[rr], &mat[ME * ((r0 + rr) * DIM + c)], xv);
  }
}

// Exchanges register bit RBIT with lane bit `lbit` across the warp: afterwards
// register bit RBIT of a thread enumerates what used to be lane bit `lbit`.
template <int RBIT, int NR, typename C>
__device__ __forceinline__ void swap_bit(C (&x)[NR], int lbit, int lane) {
  const bool hi = (lane >> lbit) & 1;
  const int mask = 1 << lbit;
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    if (i & (1 << RBIT)) continue;
    const int i0 = i, i1 = i | (1 << RBIT);
    C send = hi ? x[i0] : x[i1];
    C recv = shfl_xor_c<C>(send, mask);
    if (hi) {
      x[i0] = recv;
    } else {
      x[i1] = recv;
    }
  }
}

template <int K, int NR, typename C>
__device__ __forceinline__ void swap_all(C (&x)[NR], const int* swap_lane, int lane) {
  // Target slots are register bits 0..K-1 in the S=0 layout.
  if constexpr (K >= 1) {
    if (swap_lane[0] >= 0) swap_bit<0>(x, swap_lane[0], lane);
  }
  if constexpr (K >= 2) {
    if (swap_lane[1] >= 0) swap_bit<1>(x, swap_lane[1], lane);
  }
  if constexpr (K >= 3) {
    if (swap_lane[2] >= 0) swap_bit<2>(x, swap_lane[2], lane);
  }
  if constexpr (K >= 4) {
    if (swap_lane[3] >= 0) swap_bit<3>(x, swap_lane[3], lane);
  }
  if constexpr (K >= 5) {
    if (swap_lane[4] >= 0) swap_bit<4>(x, swap_lane[4], lane);
  }
}

// Register index layout: [GT group bits][K target slots][S low group bit].
// complex64: register bit 0 is always the 16-byte vector bit (index bit 0).
// VEC (complex64 only): lanes move 16 bytes = 2 amplitudes per access (vector
// bit = index bit 0).  Without it every lane moves one amplitude (8 or 16 bytes):
// half the registers per thread, twice the resident warps — used where the
// arithmetic, not the memory system, limits the pass (k >= 4).
template <typename real, int K, int S, int GT, bool SWAPS, bool VEC>
__global__ void __launch_bounds__(kFastThreads, K == 4 ? (VEC ? 2 : 3) : 1)
    sv_apply_fast_kernel(const __grid_constant__ FastParams<real, K> p) {
  using C = typename Cplx<real>::type;
  constexpr bool kVec = VEC;
  static_assert(!VEC || sizeof(real) == 4, "16-byte vectors of 2 amplitudes are complex64 only");
  constexpr int RB = S + K + GT;
  constexpr int NR = 1 << RB;
  constexpr int DIM = 1 << K;
  constexpr int ZB = kVec ? 6 : 5;
  constexpr int ME = MatEntry<real>::kReals;
  static_assert(!(SWAPS && S), "lane exchanges only exist in the S=0 layout");
  static_assert(kVec || S == 0, "complex128 has no vector bit");

  const int lane = threadIdx.x & 31;
  const uint64_t item =
      (uint64_t)blockIdx.x * (kFastThreads / 32) + (uint64_t)(threadIdx.x >> 5);
  if (item >= p.num_items) return;
  // The 32 lanes enumerate the 5 lowest index bits that are neither the vector
  // bit nor register-resident; with all register bits above the zone this is
  // "512 contiguous bytes per warp access", with low register bits (remap
  // mode) the lanes spread over the next free bits instead.
  constexpr int VB = kVec ? 1 : 0;
  (void)ZB;
  const uint64_t base = insert_zero_bits(((item << 5) | (uint64_t)lane) << VB, p.ins_pos, p.n_ins);
  C* __restrict__ ptr = p.state + base;

  C x[NR];
  if constexpr (kVec) {
#pragma unroll
    for (int i = 0; i < NR / 2; ++i) {
      long long off = 0;
#pragma unroll
      for (int b = 0; b < RB - 1; ++b)
        if ((i >> b) & 1) off += p.reg_off[b + 1];
      const float4 v = ldg16(ptr + off);
      x[2 * i] = make_float2(v.x, v.y);
      x[2 * i + 1] = make_float2(v.z, v.w);
    }
  } else {
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      long long off = 0;
#pragma unroll
      for (int b = 0; b < RB; ++b)
        if ((i >> b) & 1) off += p.reg_off[b];
      x[i] = ldg_elem(ptr + off);
    }
  }

  if constexpr (SWAPS) swap_all<K>(x, p.swap_lane, lane);

  // Matrix-vector product in row tiles: RT independent accumulator chains per
  // input stream keep the FMA pipe busy at 1-2 resident CTAs per SM.
  constexpr int RT = DIM < 4 ? DIM : 4;
  if constexpr (SWAPS) {
    C y[NR];
#pragma unroll
    for (int g = 0; g < (1 << GT); ++g) {
#pragma unroll
      for (int r0 = 0; r0 < DIM; r0 += RT) {
        C acc[RT];
        rows_tile<real, C, DIM, RT>(p.mat, r0, acc, [&](int c) { return x[(g << K) | c]; });
#pragma unroll
        for (int rr = 0; rr < RT; ++rr) y[(g << K) | (r0 + rr)] = acc[rr];
      }
    }
    swap_all<K>(y, p.swap_lane, lane);
    if constexpr (kVec) {
#pragma unroll
      for (int i = 0; i < NR / 2; ++i) {
        long long off = 0;
#pragma unroll
        for (int b = 0; b < RB - 1; ++b)
          if ((i >> b) & 1) off += p.reg_off[b + 1];
        stg16(ptr + off, make_float4(y[2 * i].x, y[2 * i].y, y[2 * i + 1].x, y[2 * i + 1].y));
      }
    } else {
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        long long off = 0;
#pragma unroll
        for (int b = 0; b < RB; ++b)
          if ((i >> b) & 1) off += p.reg_off[b];
        stg_elem(ptr + off, y[i]);
      }
    }
  } else if constexpr (kVec && S == 1) {
    // Vector bit is a group bit: row r of both halves forms one 16-byte store.
#pragma unroll
    for (int g = 0; g < (1 << GT); ++g) {
#pragma unroll
      for (int r0 = 0; r0 < DIM; r0 += RT) {
        C a0[RT], a1[RT];
        rows_tile<real, C, DIM, RT>(p.mat, r0, a0, [&](int c) { return x[((g << K) | c) << 1]; });
        rows_tile<real, C, DIM, RT>(p.mat, r0, a1,
                                    [&](int c) { return x[(((g << K) | c) << 1) | 1]; });
#pragma unroll
        for (int rr = 0; rr < RT; ++rr) {
          const int i = (g << K) | (r0 + rr);  // index over register bits 1..RB-1
          long long off = 0;
#pragma unroll
          for (int b = 0; b < RB - 1; ++b)
            if ((i >> b) & 1) off += p.reg_off[b + 1];
          stg16(ptr + off, make_float4(a0[rr].x, a0[rr].y, a1[rr].x, a1[rr].y));
        }
      }
    }
  } else if constexpr (kVec) {
    // S == 0, vector bit is target slot 0: rows (2i, 2i+1) form one store.
    constexpr int RT2 = 2;
#pragma unroll
    for (int g = 0; g < (1 << GT); ++g) {
#pragma unroll
      for (int r0 = 0; r0 < DIM; r0 += RT2) {
        C acc[RT2];
        rows_tile<real, C, DIM, RT2>(p.mat, r0, acc, [&](int c) { return x[(g << K) | c]; });
#pragma unroll
        for (int rr = 0; rr < RT2; rr += 2) {
          const int i = (g << (K - 1)) | ((r0 + rr) >> 1);  // register bits 1..RB-1
          long long off = 0;
#pragma unroll
          for (int b = 0; b < RB - 1; ++b)
            if ((i >> b) & 1) off += p.reg_off[b + 1];
          stg16(ptr + off, make_float4(acc[rr].x, acc[rr].y, acc[rr + 1].x, acc[rr + 1].y));
        }
      }
    }
  } else {
#pragma unroll
    for (int g = 0; g < (1 << GT); ++g) {
#pragma unroll
      for (int r0 = 0; r0 < DIM; r0 += RT) {
        C acc[RT];
        rows_tile<real, C, DIM, RT>(p.mat, r0, acc, [&](int c) { return x[(g << K) | c]; });
#pragma unroll
        for (int rr = 0; rr < RT; ++rr) {
          const int i = (g << K) | (r0 + rr);
          long long off = 0;
#pragma unroll
          for (int b = 0; b < RB; ++b)
            if ((i >> b) & 1) off += p.reg_off[b];
          stg_elem(ptr + off, acc[rr]);
        }
      }
    }
  }
}

// One thread per 2^K-amplitude group; for states too small for the zone layout.
template <typename real, int K>
__global__ void __launch_bounds__(128)
    sv_apply_small_kernel(const __grid_constant__ SmallParams<real, K> p) {
  using C = typename Cplx<real>::type;
  constexpr int DIM = 1 << K;
  constexpr int ME = MatEntry<real>::kReals;
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= p.num_groups) return;
  const uint64_t base = insert_zero_bits(g, p.tpos, K);
  C x[DIM];
#pragma unroll
  for (int j = 0; j < DIM; ++j) {
    uint64_t off = 0;
#pragma unroll
    for (int b = 0; b < K; ++b)
      if ((j >> b) & 1) off |= 1ull << p.tpos[b];
    x[j] = p.state[base | off];
  }
#pragma unroll
  for (int r = 0; r < DIM; ++r) {
    C acc = make_c<real>(0, 0);
#pragma unroll
    for (int c = 0; c < DIM; ++c)
      cmac_entry<real, C>(acc, &p.mat[ME * (r * DIM + c)], x[c]);
    uint64_t off = 0;
#pragma unroll
    for (int b = 0; b < K; ++b)
      if ((r >> b) & 1) off |= 1ull << p.tpos[b];
    p.state[base | off] = acc;
  }
}

// Any k <= 10: out[i] = sum_c M[row(i), c] * in[i with target bits := c].
struct GenericParams {
  int n;
  int k;
  int tpos[16];  // ascending
};

template <typename real>
__global__ void __launch_bounds__(256)
    sv_apply_generic_kernel(const typename Cplx<real>::type* __restrict__ in,
                            typename Cplx<real>::type* __restrict__ out,
                            const typename Cplx<real>::type* __restrict__ mat,
                            const __grid_constant__ GenericParams p) {
  using C = typename Cplx<real>::type;
  const uint64_t total = 1ull << p.n;
  const int dim = 1 << p.k;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t row = 0, mask = 0;
    for (int b = 0; b < p.k; ++b) {
      row |= ((i >> p.tpos[b]) & 1ull) << b;
      mask |= 1ull << p.tpos[b];
    }
    const uint64_t base = i & ~mask;
    C acc = make_c<real>(0, 0);
    for (int c = 0; c < dim; ++c) {
      uint64_t off = 0;
      for (int b = 0; b < p.k; ++b)
        if ((c >> b) & 1) off |= 1ull << p.tpos[b];
      const C m = mat[row * dim + c];
      cmac<real>(acc, m.x, m.y, in[base | off]);
    }
    out[i] = acc;
  }
}

// psi[i] *= diag[bits of i at targets]
struct DiagParams {
  int n;
  int k;
  int tpos[16];  // tpos[b] <-> bit b of the table index
};

template <typename real>
__global__ void __launch_bounds__(256)
    sv_apply_diag_kernel(typename Cplx<real>::type* __restrict__ state,
                         const typename Cplx<real>::type* __restrict__ diag,
                         const __grid_constant__ DiagParams p) {
  using C = typename Cplx<real>::type;
  const uint64_t total = 1ull << p.n;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t v = 0;
    for (int b = 0; b < p.k; ++b) v |= ((i >> p.tpos[b]) & 1ull) << b;
    const C d = diag[v];
    const C a = state[i];
    state[i] = cmul<real>(d.x, d.y, a);
  }
}

// Diagonal block with its table in shared memory (k <= 13 complex64 / 12
// complex128 wires: 64 KB).  A CTA walks chunks of 2^c contiguous amplitudes with
// 16-byte lane accesses; the table index of an amplitude is the OR of a part
// that depends on the chunk only (targets at or above bit c: computed once per
// chunk) and a part read from a 2^c-entry uint16 table built once per CTA
// (targets below bit c).  One read + one write of the state: HBM-bound.
constexpr int kDiagVec = 8;  // 16-byte accesses per thread per full chunk

struct DiagSmemParams {
  int n;
  int k;
  int c;         // log2 amplitudes per chunk
  int tpos[16];  // tpos[b] <-> bit b of the table index
};

template <typename real>
__global__ void __launch_bounds__(512)
    sv_apply_diag_smem_kernel(typename Cplx<real>::type* __restrict__ state,
                              const typename Cplx<real>::type* __restrict__ diag,
                              const __grid_constant__ DiagSmemParams p) {
  using C = typename Cplx<real>::type;
  extern __shared__ __align__(16) unsigned char diag_smem[];
  C* tab = reinterpret_cast<C*>(diag_smem);
  uint16_t* low_tab = reinterpret_cast<uint16_t*>(tab + (1u << p.k));
  const uint32_t chunk_elems = 1u << p.c;
  for (uint32_t i = threadIdx.x; i < (1u << p.k); i += blockDim.x) tab[i] = diag[i];
  for (uint32_t j = threadIdx.x; j < chunk_elems; j += blockDim.x) {
    uint32_t v = 0;
    for (int b = 0; b < p.k; ++b)
      if (p.tpos[b] < p.c) v |= ((j >> p.tpos[b]) & 1u) << b;
    low_tab[j] = (uint16_t)v;
  }
  __syncthreads();
  constexpr int kPer = 16 / sizeof(C);  // amplitudes per 16-byte access
  const uint64_t chunks = 1ull << (p.n - p.c);
  for (uint64_t chunk = blockIdx.x; chunk < chunks; chunk += gridDim.x) {
    const uint64_t base = chunk << p.c;
    uint32_t vh = 0;
    for (int b = 0; b < p.k; ++b)
      if (p.tpos[b] >= p.c) vh |= (uint32_t)((base >> p.tpos[b]) & 1ull) << b;
    if (chunk_elems == (uint32_t)kPer * kDiagVec * 512u) {
      // full-size chunk: all of a thread's loads are issued before the first
      // store (in-place updates otherwise serialise load -> store -> load)
      float4 x[kDiagVec];
#pragma unroll
      for (int u = 0; u < kDiagVec; ++u) {
        const uint32_t j = (threadIdx.x + u * 512u) * kPer;
        x[u] = __ldcs(reinterpret_cast<const float4*>(state + base + j));
      }
      // barrier (uniform: the chunk loop is CTA-wide): ptxas moves no memory
      // access across it; without it the later loads sink below the first stores
      // to save registers and only four stay in flight.  (A warp barrier orders
      // the accesses just as well but measured 47.6 ms per 34-qubit pass against
      // 45.4 ms: the CTA moving through its 64 KB chunk in step helps DRAM.)
      __syncthreads();
#pragma unroll
      for (int u = 0; u < kDiagVec; ++u) {
        const uint32_t j = (threadIdx.x + u * 512u) * kPer;
        float4 y;
        if constexpr (sizeof(C) == 8) {
          const C d0 = tab[vh | low_tab[j]];
          const C d1 = tab[vh | low_tab[j + 1]];
          y.x = d0.x * x[u].x - d0.y * x[u].y;
          y.y = d0.x * x[u].y + d0.y * x[u].x;
          y.z = d1.x * x[u].z - d1.y * x[u].w;
          y.w = d1.x * x[u].w + d1.y * x[u].z;
        } else {
          const C d = tab[vh | low_tab[j]];
          const double2 a = *reinterpret_cast<const double2*>(&x[u]);
          const double2 r = cmul<double>(d.x, d.y, a);
          y = *reinterpret_cast<const float4*>(&r);
        }
        __stcs(reinterpret_cast<float4*>(state + base + j), y);
      }
    } else if (chunk_elems >= (uint32_t)kPer) {
#pragma unroll 4
      for (uint32_t j = threadIdx.x * kPer; j < chunk_elems; j += blockDim.x * kPer) {
        if constexpr (sizeof(C) == 8) {
          float4 x = *reinterpret_cast<const float4*>(state + base + j);
          const C d0 = tab[vh | low_tab[j]];
          const C d1 = tab[vh | low_tab[j + 1]];
          float4 y;
          y.x = d0.x * x.x - d0.y * x.y;
          y.y = d0.x * x.y + d0.y * x.x;
          y.z = d1.x * x.z - d1.y * x.w;
          y.w = d1.x * x.w + d1.y * x.z;
          *reinterpret_cast<float4*>(state + base + j) = y;
        } else {
          const C a = state[base + j];
          const C d = tab[vh | low_tab[j]];
          state[base + j] = cmul<real>(d.x, d.y, a);
        }
      }
    } else {  // a 1-amplitude complex64 state
      if (threadIdx.x == 0) {
        const C a = state[base];
        const C d = tab[vh | low_tab[0]];
        state[base] = cmul<real>(d.x, d.y, a);
      }
    }
  }
}

// ---- host-side planning -----------------------------------------------------

struct FastPlan {
  bool feasible = false;
  int K = 0, S = 0, GT = 0;
  bool swaps = false;
  bool vec = false;
  uint64_t num_items = 0;
  int n_ins = 0;
  int ins_pos[kMaxIns] = {0};
  long long reg_off[kMaxRegBits] = {0};
  int swap_lane[kMaxRegBits] = {-1, -1, -1, -1, -1, -1};
};

// `sorted` = ascending target bit positions.  Mirrors the layout rules in the
// header comment; exported through b2q_debug_plan for the host unit tests.
// How a target bit inside the zone (a "lane target") is handled:
//   shuffle: exchanged with a spare register bit by a warp-shuffle butterfly;
//            every warp access stays 512 contiguous bytes, at the price of
//            ~8 extra instructions per amplitude pair and both x and y live.
//   remap:   it becomes an ordinary register bit and the lanes move up to the
//            next free index bits; no shuffles, but a warp access then covers
//            16-byte pieces at stride 2^(t+1) amplitudes that the thread's other
//            accesses complete to full sectors through L1/L2.
// g_lane_mode: 0 = always shuffle, 1 = always remap, 2 = per-target policy
// fitted to the B200 measurements in profiles/microbench_r1.md (default).
std::atomic<int> g_lane_mode{2};
// complex64 access width: 0 = policy (16-byte vectors, except k >= 4 blocks that
// need lane shuffles: those move one amplitude per lane, halving the registers
// that take part in the exchange), 1 = always 16-byte vectors, 2 = never.
std::atomic<int> g_vec_mode{0};


static bool remap_target(int mode, bool vec, int K, int t, bool both_low) {
  if (mode == 0) return false;
  if (mode == 1) return true;
  if (vec) {
    // complex64: sector = index bits 0-1; bits 1,2 split sectors when remapped
    if (K >= 4) return t >= 3 || !both_low;
    if (K == 3) return t >= 3;
    return false;
  }
  // complex128: sector = index bit 0
  if (K >= 4) return true;
  return t >= 3;
}

static FastPlan make_fast_plan_width(int dtype, int n, const int* sorted, int K, bool vec);

FastPlan make_fast_plan(int dtype, int n, const int* sorted, int K) {
  if (dtype != B2Q_C64) return make_fast_plan_width(dtype, n, sorted, K, false);
  const int m = g_vec_mode.load(std::memory_order_relaxed);
  if (m == 1) return make_fast_plan_width(dtype, n, sorted, K, true);
  if (m == 2) return make_fast_plan_width(dtype, n, sorted, K, false);
  FastPlan pl = make_fast_plan_width(dtype, n, sorted, K, true);
  if (K >= 4 && pl.feasible && pl.swaps) {
    const FastPlan narrow = make_fast_plan_width(dtype, n, sorted, K, false);
    if (narrow.feasible) return narrow;
  }
  return pl;
}

static FastPlan make_fast_plan_width(int dtype, int n, const int* sorted, int K, bool vec) {
  FastPlan pl;
  pl.K = K;
  pl.vec = vec;
  const int ZB = vec ? 6 : 5;
  const int VB = vec ? 1 : 0;
  const int max_k = dtype == B2Q_C64 ? 5 : 4;
  if (K < 1 || K > max_k) return pl;
  const bool vec_is_target = vec && sorted[0] == 0;
  const int mode = g_lane_mode.load(std::memory_order_relaxed);
  bool has1 = false, has2 = false;
  for (int i = 0; i < K; ++i) {
    has1 |= sorted[i] == 1;
    has2 |= sorted[i] == 2;
  }
  bool remap[16];
  int n_lane = 0;  // lane targets handled by shuffles
  for (int i = 0; i < K; ++i) {
    const bool lane_t = sorted[i] >= VB && sorted[i] < ZB;
    remap[i] = lane_t && remap_target(mode, dtype == B2Q_C64, K, sorted[i], has1 && has2);
    if (lane_t && !remap[i]) ++n_lane;
  }
  // complex64, S=0 layout: register bit 0 is physically the vector bit, so target
  // slot 0 must be index bit 0 itself or a shuffled lane target (the vector bit
  // then serves as its spare).  Keep that invariant under mixed policies.
  if (vec && !vec_is_target && n_lane > 0 && remap[0]) {
    remap[0] = false;
    ++n_lane;
  }
  pl.S = (vec && !vec_is_target && n_lane == 0) ? 1 : 0;
  pl.GT = pl.S ? std::max(0, 2 - K) : std::max(0, 3 - K);
  pl.swaps = n_lane > 0;
  auto is_target = [&](int b) {
    for (int i = 0; i < K; ++i)
      if (sorted[i] == b) return true;
    return false;
  };
  int next_extra = ZB;
  auto take_extra = [&]() {
    while (is_target(next_extra)) ++next_extra;
    return next_extra++;
  };
  std::vector<int> ins;
  const int RB = pl.S + K + pl.GT;
  for (int i = 0; i < kMaxRegBits; ++i) {
    pl.reg_off[i] = 0;
    pl.swap_lane[i] = -1;
  }
  // Target slots.
  for (int i = 0; i < K; ++i) {
    const int rbit = pl.S + i;
    const int t = sorted[i];
    if (t >= ZB || remap[i]) {
      pl.reg_off[rbit] = 1ll << t;
      ins.push_back(t);
    } else if (vec && t == 0) {
      // rbit == 0: the vector bit itself.
    } else {
      // Lane target: needs a spare register bit to trade places with.
      pl.swap_lane[i] = t - VB;
      if (vec && rbit == 0) {
        // Spare is the vector bit (index bit 0, not a target here).
      } else {
        const int e = take_extra();
        pl.reg_off[rbit] = 1ll << e;
        ins.push_back(e);
      }
    }
  }
  // Group bits on top.
  for (int g = 0; g < pl.GT; ++g) {
    const int e = take_extra();
    pl.reg_off[pl.S + K + g] = 1ll << e;
    ins.push_back(e);
  }
  (void)RB;
  std::sort(ins.begin(), ins.end());
  if ((int)ins.size() > kMaxIns) return pl;
  for (int b : ins)
    if (b >= n) return pl;
  if (n < ZB + (int)ins.size()) return pl;
  pl.n_ins = (int)ins.size();
  for (int i = 0; i < pl.n_ins; ++i) pl.ins_pos[i] = ins[i];
  // In remap mode the spare bits must not collide with the lane bits: the
  // extras were taken from >= ZB, but lanes may now reach above ZB; extras
  // are register bits (inserted positions), lanes take what is left: fine.
  pl.num_items = 1ull << (n - ZB - pl.n_ins);
  pl.feasible = true;
  return pl;
}

// Reorders a gate matrix from "first target = MSB" order to "index bit i <->
// i-th lowest target bit position", casts it to `real` and lays the entries out
// as the kernels read them (MatEntry; `packed` = false keeps plain (re, im)).
template <typename real>
void permute_matrix(const double* m128, const int* targets, const int* sorted, int K, real* out,
                    bool packed = true) {
  const int dim = 1 << K;
  // rank_of_gate_qubit[q] = index in `sorted` of targets[q]
  int rank[16];
  for (int q = 0; q < K; ++q)
    for (int i = 0; i < K; ++i)
      if (sorted[i] == targets[q]) rank[q] = i;
  std::vector<int> orig(dim);
  for (int j = 0; j < dim; ++j) {
    int idx = 0;
    for (int q = 0; q < K; ++q) {
      const int bit = (j >> rank[q]) & 1;
      idx |= bit << (K - 1 - q);
    }
    orig[j] = idx;
  }
  for (int r = 0; r < dim; ++r)
    for (int c = 0; c < dim; ++c) {
      const double* src = m128 + 2 * ((size_t)orig[r] * dim + orig[c]);
      if (packed && sizeof(real) == 4) {
        real* dst = out + 4 * ((size_t)r * dim + c);
        dst[0] = (real)src[0];
        dst[1] = (real)src[0];
        dst[2] = -(real)src[1];
        dst[3] = (real)src[1];
      } else {
        out[2 * ((size_t)r * dim + c)] = (real)src[0];
        out[2 * ((size_t)r * dim + c) + 1] = (real)src[1];
      }
    }
}

template <typename real, int K, int S, int GT, bool SWAPS, bool VEC>
int launch_fast(void* state, const FastPlan& pl, const real* mat, cudaStream_t stream) {
  FastParams<real, K> p;
  p.state = reinterpret_cast<typename Cplx<real>::type*>(state);
  p.num_items = pl.num_items;
  p.n_ins = pl.n_ins;
  for (int i = 0; i < kMaxIns; ++i) p.ins_pos[i] = pl.ins_pos[i];
  for (int i = 0; i < kMaxRegBits; ++i) {
    p.reg_off[i] = pl.reg_off[i];
    p.swap_lane[i] = pl.swap_lane[i];
  }
  std::memcpy(p.mat, mat, sizeof(real) * ((size_t)MatEntry<real>::kReals << (2 * K)));
  const uint64_t warps_per_block = kFastThreads / 32;
  const uint64_t blocks = (pl.num_items + warps_per_block - 1) / warps_per_block;
  if (blocks > 0x7fffffffull) return set_error(B2Q_ERR_UNSUPPORTED, "grid too large");
  sv_apply_fast_kernel<real, K, S, GT, SWAPS, VEC>
      <<<(unsigned)blocks, kFastThreads, 0, stream>>>(p);
  B2Q_LAUNCH_CHECK("sv_apply_fast_kernel");
  return B2Q_OK;
}

template <typename real, int K>
int dispatch_fast_k(void* state, const FastPlan& pl, const real* mat, cudaStream_t stream) {
  constexpr int GT1 = (2 - K) > 0 ? (2 - K) : 0;
  constexpr int GT0 = (3 - K) > 0 ? (3 - K) : 0;
  if constexpr (sizeof(real) == 4) {
    if (pl.vec) {
      if (pl.S == 1) return launch_fast<real, K, 1, GT1, false, true>(state, pl, mat, stream);
      if (pl.swaps) return launch_fast<real, K, 0, GT0, true, true>(state, pl, mat, stream);
      return launch_fast<real, K, 0, GT0, false, true>(state, pl, mat, stream);
    }
  }
  if (pl.swaps) return launch_fast<real, K, 0, GT0, true, false>(state, pl, mat, stream);
  return launch_fast<real, K, 0, GT0, false, false>(state, pl, mat, stream);
}

template <typename real, int K>
int launch_small(void* state, int n, const int* sorted, const real* mat, cudaStream_t stream) {
  SmallParams<real, K> p;
  p.state = reinterpret_cast<typename Cplx<real>::type*>(state);
  p.num_groups = 1ull << (n - K);
  for (int i = 0; i < K; ++i) p.tpos[i] = sorted[i];
  std::memcpy(p.mat, mat, sizeof(real) * ((size_t)MatEntry<real>::kReals << (2 * K)));
  const uint64_t blocks = (p.num_groups + 127) / 128;
  if (blocks > 0x7fffffffull) return set_error(B2Q_ERR_UNSUPPORTED, "grid too large");
  sv_apply_small_kernel<real, K><<<(unsigned)blocks, 128, 0, stream>>>(p);
  B2Q_LAUNCH_CHECK("sv_apply_small_kernel");
  return B2Q_OK;
}

template <typename real>
int apply_generic(void* state, int n, const int* sorted, int K, const real* mat, void* scratch,
                  cudaStream_t stream) {
  using C = typename Cplx<real>::type;
  if (scratch == nullptr)
    return set_error(B2Q_ERR_INVALID,
                     "a %d-qubit gate needs the out-of-place kernel: pass a scratch buffer", K);
  const size_t dim = (size_t)1 << K;
  C* dmat = nullptr;
  keep_async_pool_memory();
  B2Q_CUDA_CHECK(cudaMallocAsync((void**)&dmat, sizeof(C) * dim * dim, stream));
  B2Q_CUDA_CHECK(
      cudaMemcpyAsync(dmat, mat, sizeof(C) * dim * dim, cudaMemcpyHostToDevice, stream));
  // The host matrix buffer is reused by the caller: make the copy complete.
  B2Q_CUDA_CHECK(cudaStreamSynchronize(stream));
  GenericParams p;
  p.n = n;
  p.k = K;
  for (int i = 0; i < K; ++i) p.tpos[i] = sorted[i];
  const uint64_t total = 1ull << n;
  const uint64_t blocks = std::min<uint64_t>((total + 255) / 256, 148ull * 64);
  sv_apply_generic_kernel<real><<<(unsigned)blocks, 256, 0, stream>>>(
      reinterpret_cast<const C*>(state), reinterpret_cast<C*>(scratch), dmat, p);
  B2Q_LAUNCH_CHECK("sv_apply_generic_kernel");
  B2Q_CUDA_CHECK(
      cudaMemcpyAsync(state, scratch, sizeof(C) * total, cudaMemcpyDeviceToDevice, stream));
  B2Q_CUDA_CHECK(cudaFreeAsync(dmat, stream));
  return B2Q_OK;
}

// b2q_apply_tc.cu
bool tc_applicable(int dtype, int n, int K);
int launch_tc(void* state, int n, int K, const int* sorted, const float* mat, cudaStream_t stream);
int launch_tc_exchange(void* state, int n, int K, const int* sorted, const float* mat,
                       void* out_local, void* out_peer, int xbit, int gval, cudaStream_t stream);

bool tile_bits_for(int n, int nb, const int* ks, const int* targets, int* tbits);
int launch_tc_tile(void* state, int n, int nb, const int (*sorted)[5], const int* tbits,
                   const float* const* mats, cudaStream_t stream);

// Widens block (m128 on targets[0..k)) to 5 targets inside the tile bits (identity
// on the added wires, which become the most significant ones) and brings it to
// sorted-target order: `sorted5` and `plain` (32 x 32 (re, im) float pairs).
static void widen_block_to_5(const double* m128, const int* targets, int k, const int* tbits,
                             int* sorted5, float* plain) {
  int wide[5];
  const int extra = 5 - k;
  int e = 0;
  for (int i = 12; i >= 0 && e < extra; --i) {
    bool used = false;
    for (int q = 0; q < k; ++q) used |= targets[q] == tbits[i];
    if (!used) wide[e++] = tbits[i];
  }
  for (int q = 0; q < k; ++q) wide[extra + q] = targets[q];
  const int dk = 1 << k;
  std::vector<double> big((size_t)2 * 32 * 32, 0.0);
  for (int hi = 0; hi < (1 << extra); ++hi)
    for (int r = 0; r < dk; ++r)
      for (int c = 0; c < dk; ++c) {
        const size_t dst = 2 * ((size_t)(hi * dk + r) * 32 + (size_t)(hi * dk + c));
        big[dst] = m128[2 * ((size_t)r * dk + c)];
        big[dst + 1] = m128[2 * ((size_t)r * dk + c) + 1];
      }
  for (int i = 0; i < 5; ++i) sorted5[i] = wide[i];
  std::sort(sorted5, sorted5 + 5);
  permute_matrix<float>(big.data(), wide, sorted5, 5, plain, /*packed=*/false);
}

template <typename real>
int apply_matrix_t(void* state, int dtype, int n, const double* m128, const int* targets, int K,
                   void* scratch, cudaStream_t stream) {
  int sorted[16];
  for (int i = 0; i < K; ++i) sorted[i] = targets[i];
  std::sort(sorted, sorted + K);
  for (int i = 0; i < K; ++i) {
    B2Q_REQUIRE(sorted[i] >= 0 && sorted[i] < n, "target bit %d out of range for %d qubits",
                sorted[i], n);
    B2Q_REQUIRE(i == 0 || sorted[i] != sorted[i - 1], "duplicate target bit %d", sorted[i]);
  }
  if constexpr (sizeof(real) == 4) {
    if (tc_applicable(dtype, n, K)) {
      std::vector<float> plain((size_t)2 << (2 * K));
      permute_matrix<float>(m128, targets, sorted, K, plain.data(), /*packed=*/false);
      return launch_tc(state, n, K, sorted, plain.data(), stream);
    }
  }
  std::vector<real> mat((size_t)MatEntry<real>::kReals << (2 * K));
  permute_matrix<real>(m128, targets, sorted, K, mat.data());
  const FastPlan pl = make_fast_plan(dtype, n, sorted, K);
  if (pl.feasible) {
    switch (K) {
      case 1: return dispatch_fast_k<real, 1>(state, pl, mat.data(), stream);
      case 2: return dispatch_fast_k<real, 2>(state, pl, mat.data(), stream);
      case 3: return dispatch_fast_k<real, 3>(state, pl, mat.data(), stream);
      case 4: return dispatch_fast_k<real, 4>(state, pl, mat.data(), stream);
      case 5:
        if constexpr (sizeof(real) == 4)
          return dispatch_fast_k<real, 5>(state, pl, mat.data(), stream);
        break;
    }
  }
  const int max_small = sizeof(real) == 4 ? 5 : 4;
  if (K <= max_small && n <= 20) {
    switch (K) {
      case 1: return launch_small<real, 1>(state, n, sorted, mat.data(), stream);
      case 2: return launch_small<real, 2>(state, n, sorted, mat.data(), stream);
      case 3: return launch_small<real, 3>(state, n, sorted, mat.data(), stream);
      case 4: return launch_small<real, 4>(state, n, sorted, mat.data(), stream);
      case 5:
        if constexpr (sizeof(real) == 4)
          return launch_small<real, 5>(state, n, sorted, mat.data(), stream);
        break;
    }
  }
  std::vector<real> plain((size_t)2 << (2 * K));
  permute_matrix<real>(m128, targets, sorted, K, plain.data(), /*packed=*/false);
  return apply_generic<real>(state, n, sorted, K, plain.data(), scratch, stream);
}

}  // namespace b2q

using namespace b2q;

extern "C" int b2q_sv_apply_matrix(void* state, int dtype, int n_qubits,
                                   const double* matrix_c128, const int* targets, int k,
                                   void* scratch, void* stream) {
  B2Q_REQUIRE(state != nullptr && matrix_c128 != nullptr && targets != nullptr, "null argument");
  B2Q_REQUIRE(dtype == B2Q_C64 || dtype == B2Q_C128, "bad dtype %d", dtype);
  B2Q_REQUIRE(n_qubits >= 1 && n_qubits <= 40, "n_qubits=%d out of range", n_qubits);
  B2Q_REQUIRE(k >= 1 && k <= 10 && k <= n_qubits, "k=%d out of range (n=%d)", k, n_qubits);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == B2Q_C64)
    return apply_matrix_t<float>(state, dtype, n_qubits, matrix_c128, targets, k, scratch, s);
  return apply_matrix_t<double>(state, dtype, n_qubits, matrix_c128, targets, k, scratch, s);
}

extern "C" int b2q_sv_apply_batch(void* state, int dtype, int n_qubits, int num_gates,
                                  const int* ks, const int* targets, const double* matrices_c128,
                                  void* scratch, void* stream) {
  B2Q_REQUIRE(num_gates >= 0, "num_gates < 0");
  size_t toff = 0, moff = 0;
  for (int g = 0; g < num_gates; ++g) {
    const int k = ks[g];
    const int rc = b2q_sv_apply_matrix(state, dtype, n_qubits, matrices_c128 + moff,
                                       targets + toff, k, scratch, stream);
    if (rc != B2Q_OK) return rc;
    toff += k;
    moff += (size_t)2 << (2 * k);
  }
  return B2Q_OK;
}

extern "C" int b2q_tile_blocks_feasible(int dtype, int n_qubits, int num_blocks, const int* ks,
                                        const int* targets) {
  if (dtype != B2Q_C64 || ks == nullptr || targets == nullptr) return 0;
  if (!tc_applicable(dtype, n_qubits, 5)) return 0;
  int tbits[16];
  return tile_bits_for(n_qubits, num_blocks, ks, targets, tbits) ? 1 : 0;
}

extern "C" int b2q_sv_apply_tile_blocks(void* state, int dtype, int n_qubits, int num_blocks,
                                        const int* ks, const int* targets,
                                        const double* matrices_c128, void* stream) {
  B2Q_REQUIRE(state != nullptr && ks != nullptr && targets != nullptr && matrices_c128 != nullptr,
              "null argument");
  B2Q_REQUIRE(dtype == B2Q_C64, "tile passes are complex64 only");
  B2Q_REQUIRE(num_blocks >= 1 && num_blocks <= 2, "a tile pass takes 1 or 2 blocks, got %d", num_blocks);
  B2Q_REQUIRE(tc_applicable(dtype, n_qubits, 5), "tensor-core kernels not applicable (n=%d)", n_qubits);
  int tbits[16];
  B2Q_REQUIRE(tile_bits_for(n_qubits, num_blocks, ks, targets, tbits),
              "blocks do not fit one 13-bit tile (or bad targets)");
  int sorted[2][5];
  std::vector<float> plain[2];
  const float* mats[2] = {nullptr, nullptr};
  size_t toff = 0, moff = 0;
  for (int b = 0; b < num_blocks; ++b) {
    const int k = ks[b];
    for (int i = 0; i < k; ++i)
      for (int j = 0; j < i; ++j)
        B2Q_REQUIRE(targets[toff + i] != targets[toff + j], "duplicate target bit %d", targets[toff + i]);
    plain[b].resize((size_t)2 * 32 * 32);
    widen_block_to_5(matrices_c128 + moff, targets + toff, k, tbits, sorted[b], plain[b].data());
    mats[b] = plain[b].data();
    toff += k;
    moff += (size_t)2 << (2 * k);
  }
  return launch_tc_tile(state, n_qubits, num_blocks, sorted, tbits, mats,
                        reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int b2q_sv_apply_diagonal(void* state, int dtype, int n_qubits,
                                     const double* diag_c128, const int* targets, int k,
                                     void* stream) {
  B2Q_REQUIRE(state != nullptr && diag_c128 != nullptr && targets != nullptr, "null argument");
  B2Q_REQUIRE(dtype == B2Q_C64 || dtype == B2Q_C128, "bad dtype %d", dtype);
  B2Q_REQUIRE(k >= 1 && k <= 16 && k <= n_qubits, "k=%d out of range", k);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  DiagParams p;
  p.n = n_qubits;
  p.k = k;
  // table index: first target = MSB  ->  bit (k-1-q) of the index <-> targets[q]
  for (int q = 0; q < k; ++q) {
    B2Q_REQUIRE(targets[q] >= 0 && targets[q] < n_qubits, "target bit out of range");
    p.tpos[k - 1 - q] = targets[q];
  }
  // The kernels want index bit b <-> the b-th LOWEST target (neighbouring
  // amplitudes then read neighbouring table entries: no shared-memory bank
  // conflicts, cache-line locality for the global table); other orders get their
  // table permuted here.
  std::vector<double> permuted;
  {
    bool ascending = true;
    for (int b = 1; b < k; ++b) ascending = ascending && p.tpos[b] > p.tpos[b - 1];
    if (!ascending) {
      int sorted[16], from_bit[16];
      for (int b = 0; b < k; ++b) sorted[b] = p.tpos[b];
      std::sort(sorted, sorted + k);
      for (int b = 1; b < k; ++b)
        B2Q_REQUIRE(sorted[b] != sorted[b - 1], "duplicate target bit %d", sorted[b]);
      for (int b = 0; b < k; ++b)
        for (int o = 0; o < k; ++o)
          if (p.tpos[o] == sorted[b]) from_bit[b] = o;  // new index bit b = old index bit o
      const size_t dim0 = (size_t)1 << k;
      permuted.resize(2 * dim0);
      for (size_t v = 0; v < dim0; ++v) {
        size_t old = 0;
        for (int b = 0; b < k; ++b)
          if ((v >> b) & 1) old |= (size_t)1 << from_bit[b];
        permuted[2 * v] = diag_c128[2 * old];
        permuted[2 * v + 1] = diag_c128[2 * old + 1];
      }
      diag_c128 = permuted.data();
      for (int b = 0; b < k; ++b) p.tpos[b] = sorted[b];
    }
  }
  const size_t dim = (size_t)1 << k;
  const uint64_t total = 1ull << n_qubits;
  const uint64_t blocks = std::min<uint64_t>((total + 255) / 256, 148ull * 64);
  const size_t esize = dtype == B2Q_C64 ? sizeof(float2) : sizeof(double2);
  // table in shared memory when it fits 64 KB (k <= 13 complex64, 12 complex128)
  const bool smem_path = dim * esize <= (64u << 10);
  DiagSmemParams sp;
  size_t smem_bytes = 0;
  unsigned smem_grid = 1;
  if (smem_path) {
    sp.n = n_qubits;
    sp.k = k;
    sp.c = std::min(n_qubits, dtype == B2Q_C64 ? 13 : 12);  // 4096 16-byte vectors per chunk
    for (int b = 0; b < 16; ++b) sp.tpos[b] = p.tpos[b];
    smem_bytes = dim * esize + (sizeof(uint16_t) << sp.c);
    smem_grid = (unsigned)std::min<uint64_t>(1ull << (n_qubits - sp.c), 148ull * 2);
    static bool attr_set[64][2] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev][dtype == B2Q_C64 ? 0 : 1]) {
      if (dtype == B2Q_C64)
        B2Q_CUDA_CHECK(cudaFuncSetAttribute(sv_apply_diag_smem_kernel<float>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, 96 << 10));
      else
        B2Q_CUDA_CHECK(cudaFuncSetAttribute(sv_apply_diag_smem_kernel<double>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, 96 << 10));
      attr_set[dev][dtype == B2Q_C64 ? 0 : 1] = true;
    }
  }
  void* ddiag = nullptr;
  keep_async_pool_memory();
  if (dtype == B2Q_C64) {
    std::vector<float> h(2 * dim);
    for (size_t i = 0; i < 2 * dim; ++i) h[i] = (float)diag_c128[i];
    B2Q_CUDA_CHECK(cudaMallocAsync(&ddiag, sizeof(float2) * dim, s));
    // (a copy from pageable memory returns once the source has been staged)
    B2Q_CUDA_CHECK(cudaMemcpyAsync(ddiag, h.data(), sizeof(float2) * dim, cudaMemcpyHostToDevice, s));
    if (smem_path)
      sv_apply_diag_smem_kernel<float><<<smem_grid, 512, smem_bytes, s>>>(
          reinterpret_cast<float2*>(state), reinterpret_cast<const float2*>(ddiag), sp);
    else
      sv_apply_diag_kernel<float><<<(unsigned)blocks, 256, 0, s>>>(
          reinterpret_cast<float2*>(state), reinterpret_cast<const float2*>(ddiag), p);
  } else {
    B2Q_CUDA_CHECK(cudaMallocAsync(&ddiag, sizeof(double2) * dim, s));
    B2Q_CUDA_CHECK(
        cudaMemcpyAsync(ddiag, diag_c128, sizeof(double2) * dim, cudaMemcpyHostToDevice, s));
    if (smem_path)
      sv_apply_diag_smem_kernel<double><<<smem_grid, 512, smem_bytes, s>>>(
          reinterpret_cast<double2*>(state), reinterpret_cast<const double2*>(ddiag), sp);
    else
      sv_apply_diag_kernel<double><<<(unsigned)blocks, 256, 0, s>>>(
          reinterpret_cast<double2*>(state), reinterpret_cast<const double2*>(ddiag), p);
  }
  B2Q_LAUNCH_CHECK("sv_apply_diag_kernel");
  B2Q_CUDA_CHECK(cudaFreeAsync(ddiag, s));
  return B2Q_OK;
}

extern "C" int b2q_set_vec_mode(int mode) {
  B2Q_REQUIRE(mode >= 0 && mode <= 2, "vec mode must be 0, 1 or 2");
  g_vec_mode.store(mode, std::memory_order_relaxed);
  return B2Q_OK;
}

extern "C" int b2q_set_lane_mode(int mode) {
  B2Q_REQUIRE(mode >= 0 && mode <= 2, "lane mode must be 0, 1 or 2");
  g_lane_mode.store(mode, std::memory_order_relaxed);
  return B2Q_OK;
}

// Host-only: exposes the fast-path plan for unit tests (no GPU needed).
// out[0]=feasible, [1]=S, [2]=GT, [3]=swaps, [4]=n_ins, [5..10]=ins_pos,
// [11..16]=log2(reg_off) or -1, [17..22]=swap_lane, [23]=log2(num_items).
extern "C" int b2q_dist_apply_exchange(const void* shard_in, void* out_local, void* out_peer,
                                       int dtype, int n_local, const double* matrix_c128,
                                       const int* targets, int k, int local_bit,
                                       int my_global_bit_value, void* stream) {
  B2Q_REQUIRE(shard_in != nullptr && out_local != nullptr && out_peer != nullptr &&
                  matrix_c128 != nullptr && targets != nullptr,
              "null argument");
  B2Q_REQUIRE(shard_in != out_local, "the fused exchange is out of place");
  B2Q_REQUIRE(dtype == B2Q_C64, "fused apply+exchange is complex64 only");
  B2Q_REQUIRE(k == 4 || k == 5, "fused apply+exchange needs a 4- or 5-qubit block, got %d", k);
  B2Q_REQUIRE(n_local >= k + 7, "shard too small: %d local qubits", n_local);
  B2Q_REQUIRE(local_bit >= 1 && local_bit < n_local, "local bit %d must be in [1, %d)", local_bit,
              n_local);
  B2Q_REQUIRE(my_global_bit_value == 0 || my_global_bit_value == 1, "bad global bit value");
  int sorted[8];
  for (int i = 0; i < k; ++i) sorted[i] = targets[i];
  std::sort(sorted, sorted + k);
  for (int i = 0; i < k; ++i) {
    B2Q_REQUIRE(sorted[i] >= 0 && sorted[i] < n_local, "target bit %d out of range", sorted[i]);
    B2Q_REQUIRE(i == 0 || sorted[i] != sorted[i - 1], "duplicate target bit %d", sorted[i]);
  }
  std::vector<float> plain((size_t)2 << (2 * k));
  permute_matrix<float>(matrix_c128, targets, sorted, k, plain.data(), /*packed=*/false);
  return launch_tc_exchange(const_cast<void*>(shard_in), n_local, k, sorted, plain.data(), out_local,
                            out_peer, local_bit, my_global_bit_value,
                            reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int b2q_debug_plan(int dtype, int n_qubits, const int* targets, int k, int* out) {
  int sorted[16];
  for (int i = 0; i < k; ++i) sorted[i] = targets[i];
  std::sort(sorted, sorted + k);
  const FastPlan pl = make_fast_plan(dtype, n_qubits, sorted, k);
  out[0] = pl.feasible;
  out[1] = pl.S;
  out[2] = pl.GT;
  out[3] = (pl.swaps ? 1 : 0) | (pl.vec ? 2 : 0);
  out[4] = pl.n_ins;
  for (int i = 0; i < 6; ++i) out[5 + i] = pl.ins_pos[i];
  for (int i = 0; i < 6; ++i) {
    int lg = -1;
    if (pl.reg_off[i] > 0) {
      lg = 0;
      while ((1ll << lg) < pl.reg_off[i]) ++lg;
    }
    out[11 + i] = lg;
  }
  for (int i = 0; i < 6; ++i) out[17 + i] = pl.swap_lane[i];
  int lg = 0;
  while ((1ull << lg) < pl.num_items) ++lg;
  out[23] = pl.feasible ? lg : -1;
  return B2Q_OK;
}

// Host-only: the matrix permutation used by every kernel, for unit tests.
extern "C" int b2q_debug_permute_matrix(const double* m128, const int* targets, int k,
                                        double* out) {
  int sorted[16];
  for (int i = 0; i < k; ++i) sorted[i] = targets[i];
  std::sort(sorted, sorted + k);
  permute_matrix<double>(m128, targets, sorted, k, out, /*packed=*/false);
  return B2Q_OK;
}
