// 5- and 6-qubit fused blocks on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// A dense k-qubit block costs 4 * 2^k real FMAs per amplitude; at k = 5 that
// is 128, past the FP32 CUDA-core roofline of a B200, so the
// register kernel of b2q_apply.cu runs at ~0.25 of the HBM roofline.  Here the
// multiply runs on tcgen05.mma instead while the pass stays one in-place
// streaming sweep over HBM:
//
//   * one CTA = 128 threads = 128 amplitude groups = the M dimension of one
//     UMMA tile; thread t owns group t (32 amplitudes = 64 reals, loaded with
//     8-byte accesses that are contiguous across the warp),
//   * the complex product is embedded in a real GEMM:  D[128 x 64] =
//     A[128 x 64] * B^T, A row = (re0, im0, re1, im1, ...), B built on the host
//     from the gate matrix ([[Mr, -Mi], [Mi, Mr]] interleaved),
//   * fp32 accuracy on TF32 tensor cores by the 3xTF32 split: a = a_hi + a_lo,
//     b = b_hi + b_lo, D = a_hi b_hi + a_lo b_hi + a_hi b_lo (fp32 accumulate in
//     TMEM); the dropped a_lo b_lo term is ~2^-22 relative,
//   * A never touches shared memory: each thread writes its row (hi and lo)
//     straight from registers into TMEM (tcgen05.st), B (hi, lo; 32 KB) sits in
//     shared memory in the UMMA K-major canonical layout for the whole life of
//     the persistent CTA, D comes back with tcgen05.ld and is stored in place.
//
// Replaces the same reference call sites as b2q_apply.cu
// (linalg/transformations.py:105-172 for 5-qubit matrices).
#include "b2q_common.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

namespace b2q {

constexpr int kTcThreads = 128;
constexpr int kTcMaxK = 6;

// K = 5: 32 complex = 64 reals per group, TMEM 256 columns, 32 KB of B, 2 CTAs/SM.
// K = 6: 64 complex = 128 reals per group, TMEM 512 columns, 128 KB of B, 1 CTA/SM.
template <int K>
struct TcTraits {
  static constexpr int kDim = 1 << K;     // complex amplitudes per group
  static constexpr int kN = 2 * kDim;     // reals: N and K of the real GEMM
  static constexpr int kCols = 4 * kN;    // TMEM columns: A_hi | A_lo | D0 | D1
  static constexpr size_t kBBytes = 2ull * kN * kN * sizeof(float);  // B_hi + B_lo
  static constexpr int kMinBlocks = K == 4 ? 4 : (K == 5 ? 2 : 1);
};

struct TcParams {
  float2* state;
  uint64_t num_tiles;  // groups / 128
  int tpos[kTcMaxK];   // ascending target positions
  const float* bmat;   // device: B_hi then B_lo, UMMA K-major no-swizzle layout
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Round to the TF32 grid (10 explicit mantissa bits), ties away from zero: two
// integer instructions.  cvt.rna.tf32.f32 compiles to ~5 (NaN/Inf handling the
// amplitudes never need).
__device__ __forceinline__ uint32_t to_tf32(float x) {
  return (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, "
      "%12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),
        "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]),
        "r"(v[15])
      : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, "
      "%12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// UMMA shared-memory descriptor, K-major, no swizzle (version 1 = Blackwell):
// [0,14) start>>4, [16,30) leading byte offset>>4 (between the two 16-byte K
// chunks of one MMA), [32,46) stride byte offset>>4 (between 8-row groups).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
  d |= 1ull << 46;
  return d;
}

__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}\n"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u),
        "r"(0u), "r"(0u)
      : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}

// LOW selects the HBM access pattern for targets on the lowest index bits, so
// that every warp request still covers whole 32-byte sectors:
//   0: no target on index bit 0 (and bit 1 free or bit 0/1 both free): 8-byte
//      accesses, consecutive threads = consecutive amplitudes;
//   1: index bit 0 is a target: members (2i, 2i+1) of a group are adjacent, one
//      16-byte access per member pair;
//   2: index bit 1 is a target and bit 0 is not: groups (g, g+1) interleave at
//      8 bytes; the even thread of a pair fetches member 2i of both groups, the
//      odd thread member 2i+1 (16 bytes each), and they trade halves by shuffle.
template <int K, int LOW>
__global__ void __launch_bounds__(kTcThreads, TcTraits<K>::kMinBlocks)
    sv_apply_tc_kernel(const __grid_constant__ TcParams p) {
  constexpr int kTcK = K;
  constexpr int kTcDim = TcTraits<K>::kDim;
  constexpr int kTcN = TcTraits<K>::kN;
  constexpr int kTcCols = TcTraits<K>::kCols;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* sB = reinterpret_cast<float*>(smem_raw);  // B_hi then B_lo
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  // element offset of member j of a group (same for every tile): one LDS instead
  // of ~10 integer instructions per access
  __shared__ __align__(8) uint64_t s_off[TcTraits<K>::kDim];

  const int tid = threadIdx.x;
  const int warp = tid >> 5;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_s)),
                 "r"((uint32_t)kTcCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int j = tid; j < kTcDim; j += kTcThreads) {
    uint64_t off = 0;
    for (int b = 0; b < kTcK; ++b)
      if ((j >> b) & 1) off += 1ull << p.tpos[b];
    s_off[j] = off;
  }
  {
    const float4* src = reinterpret_cast<const float4*>(p.bmat);
    float4* dst = reinterpret_cast<float4*>(sB);
    for (int i = tid; i < 2 * kTcN * kTcN / 4; i += kTcThreads) dst[i] = src[i];
  }
  // generic-proxy writes of B must be visible to the tensor core (async proxy)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
  const uint32_t d_col = 2 * kTcN;  // D behind A_hi and A_lo

  // instruction descriptor: D=F32, A=B=TF32, both K-major, N=kTcN, M=128
  constexpr uint32_t idesc =
      (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcN >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t sb_hi = smem_u32(sB);
  const uint32_t sb_lo = smem_u32(sB + kTcN * kTcN);
  constexpr uint32_t kLbo = kTcN * 16;  // bytes between consecutive 16-byte K chunks
  constexpr uint32_t kSbo = 128;        // bytes between 8-row groups

  // Software pipeline: the loads of tile i+1 are issued as soon as tile i's
  // amplitudes have been handed to TMEM, so they are in flight while the tensor
  // core multiplies and the epilogue stores tile i.
  uint32_t parity = 0;
  uint64_t tile = blockIdx.x;
  float2 x[kTcDim];
  float2* ptr = p.state;
  const int odd = tid & 1;
  auto issue_loads = [&](uint64_t t) {
    const uint64_t g = t * kTcThreads + (uint64_t)tid;
    if constexpr (LOW == 0) {
      ptr = p.state + insert_zero_bits(g, p.tpos, kTcK);
#pragma unroll
      for (int j = 0; j < kTcDim; ++j) {
        const float2* q = ptr + s_off[j];
        asm volatile("ld.global.v2.f32 {%0,%1}, [%2];" : "=f"(x[j].x), "=f"(x[j].y) : "l"(q));
      }
    } else {
      // LOW == 2: `ptr` is the base of the thread PAIR (group g & ~1)
      ptr = p.state + insert_zero_bits(LOW == 2 ? (g & ~1ull) : g, p.tpos, kTcK);
#pragma unroll
      for (int i = 0; i < kTcDim / 2; ++i) {
        const float2* q = ptr + s_off[2 * i + (LOW == 2 ? odd : 0)];
        asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(x[2 * i].x), "=f"(x[2 * i].y), "=f"(x[2 * i + 1].x), "=f"(x[2 * i + 1].y)
                     : "l"(q));
      }
    }
  };
  if (tile < p.num_tiles) issue_loads(tile);
  while (tile < p.num_tiles) {
    float2* const cur = ptr;
    if constexpr (LOW == 2) {
      // raw: x[2i] = member (2i + odd) of group g & ~1, x[2i+1] = same member of
      // the next group.  Keep the half that is ours, trade the other.
#pragma unroll
      for (int i = 0; i < kTcDim / 2; ++i) {
        const float2 keep = odd ? x[2 * i + 1] : x[2 * i];
        const float2 send = odd ? x[2 * i] : x[2 * i + 1];
        float2 recv;
        recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
        recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
        x[2 * i] = odd ? recv : keep;
        x[2 * i + 1] = odd ? keep : recv;
      }
    }
    // A row -> TMEM, 16 columns at a time: hi at [0,N), lo at [N,2N)
#pragma unroll
    for (int c16 = 0; c16 < kTcN / 16; ++c16) {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int col = c16 * 16 + e;
        const float a = (col & 1) ? x[col >> 1].y : x[col >> 1].x;
        const uint32_t h = to_tf32(a);
        hi[e] = h;
        lo[e] = to_tf32(a - __uint_as_float(h));  // rounded, not truncated by the MMA
      }
      tmem_st16(lane_base + (uint32_t)(c16 * 16), hi);
      tmem_st16(lane_base + (uint32_t)(kTcN + c16 * 16), lo);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    const uint64_t next = tile + gridDim.x;
    if (next < p.num_tiles) issue_loads(next);
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // The tensor core truncates (rounds toward zero) every time it adds an MMA
      // result into the fp32 accumulator, which shrinks the state by ~3e-8 per
      // accumulation of a full-size partial sum.  Keep those few: the two small
      // cross terms (a_lo b_hi, a_hi b_lo: ~2^-11 of the result) go first, while
      // the accumulator is still tiny, and the main product a_hi b_hi is split
      // over two accumulators (4 K-steps each) that are added in fp32
      // round-to-nearest by the CUDA cores in the epilogue.
      uint32_t acc0 = 0;
#pragma unroll
      for (int prod = 0; prod < 2; ++prod) {
        const uint32_t a_col = (prod == 0) ? (uint32_t)kTcN : 0u;  // A_lo, then A_hi
        const uint32_t sb = (prod == 0) ? sb_hi : sb_lo;            // B_hi, then B_lo
#pragma unroll
        for (int ks = 0; ks < kTcN / 8; ++ks) {
          const uint64_t bd = umma_desc(sb + (uint32_t)(ks * 2) * kLbo, kLbo, kSbo);
          umma_tf32_ts(tmem_base + d_col, tmem_base + a_col + (uint32_t)(ks * 8), bd, idesc, acc0);
          acc0 = 1;
        }
      }
#pragma unroll
      for (int ks = 0; ks < kTcN / 8; ++ks) {
        const uint64_t bd = umma_desc(sb_hi + (uint32_t)(ks * 2) * kLbo, kLbo, kSbo);
        const bool second = ks >= kTcN / 16;
        umma_tf32_ts(tmem_base + d_col + (second ? (uint32_t)kTcN : 0u),
                     tmem_base + (uint32_t)(ks * 8), bd, idesc,
                     (second && ks == kTcN / 16) ? 0u : 1u);
      }
      asm volatile(
          "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
              smem_u32(&mbar))
          : "memory");
    }
    mbar_wait(smem_u32(&mbar), parity);
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // D0 + D1 -> registers -> HBM (in place)
#pragma unroll
    for (int c16 = 0; c16 < kTcN / 16; ++c16) {
      uint32_t d[16], d1[16];
      tmem_ld16(lane_base + d_col + (uint32_t)(c16 * 16), d);
      tmem_ld16(lane_base + d_col + (uint32_t)(kTcN + c16 * 16), d1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if constexpr (LOW == 0) {
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          const int r = (c16 * 16 + e) >> 1;
          float2* q = cur + s_off[r];
          const float re = __uint_as_float(d[e]) + __uint_as_float(d1[e]);
          const float im = __uint_as_float(d[e + 1]) + __uint_as_float(d1[e + 1]);
          asm volatile("st.global.v2.f32 [%0], {%1,%2};" ::"l"(q), "f"(re), "f"(im) : "memory");
        }
      } else {
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          const int r = (c16 * 16 + e) >> 1;  // even member; r + 1 is its neighbour
          float2 o0, o1;
          o0.x = __uint_as_float(d[e]) + __uint_as_float(d1[e]);
          o0.y = __uint_as_float(d[e + 1]) + __uint_as_float(d1[e + 1]);
          o1.x = __uint_as_float(d[e + 2]) + __uint_as_float(d1[e + 2]);
          o1.y = __uint_as_float(d[e + 3]) + __uint_as_float(d1[e + 3]);
          if constexpr (LOW == 2) {
            // even thread writes member r of both groups, odd thread member r + 1
            const float2 send = odd ? o0 : o1;
            float2 recv;
            recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
            recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
            const float2 first = odd ? recv : o0;
            const float2 second = odd ? o1 : recv;
            o0 = first;
            o1 = second;
          }
          float2* q = cur + s_off[r + (LOW == 2 ? odd : 0)];
          asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(q), "f"(o0.x), "f"(o0.y),
                       "f"(o1.x), "f"(o1.y)
                       : "memory");
        }
      }
    }
    // all reads of D must finish before the next tile's MMAs overwrite it
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    tile = next;
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)kTcCols)
                 : "memory");
  }
}

// ---- staged variant: coalesced for ANY target positions ----------------------
//
// With targets on the lowest index bits the amplitudes of one group sit next to
// each other in memory, so "thread = group" addressing makes every warp request
// touch 16-32 different cache lines (4.9 ms per pass for targets 0-4).  Here a
// warp instead moves its REGION — 32 groups x 2^K members = the index bits
// {targets} U {5 lowest non-target bits}, which always contains index bits 0-4 —
// between HBM and shared memory with 16-byte lane accesses that are contiguous
// over 256-512 bytes (cp.async, double buffered, no registers involved), and
// the per-group gather / scatter happens on shared memory.  An XOR swizzle of
// local-index bits 1-3, derived on the host from where the group bits fall,
// makes both access patterns bank-conflict free.
constexpr int kStageMaxK = 5;

struct TcStagedParams {
  float2* state;
  uint64_t num_tiles;  // CTA tiles (4 warp regions each)
  const float* bmat;
  int rpos[kStageMaxK + 5];  // ascending index positions of the region bits
  int p5;                    // rpos[5]: index position behind lane bit 4 of a request
  int gpos[5];               // local positions of the 5 group bits (lane bit i)
  int swz_src[3];            // local bit swz_src[i] (>= 4, or 20 = unused) is XORed ...
  int swz_dst[3];            // ... into local bit swz_dst[i] (1..3)
  // Results go to out_local (== state for an in-place pass).  With xbit >= 0 the
  // pass is fused with a global<->local qubit exchange (b2q_dist_apply_exchange):
  // an amplitude whose index bit `xbit` differs from this rank's global bit value
  // `gval` is stored into the PARTNER's buffer out_peer (peer memory over
  // NVLink), with that bit set to gval; the others stay in out_local.
  float2* out_local;
  float2* out_peer;
  int xbit;
  int gval;
  int early;      // 1: the next region's copy is issued at the top of the iteration
  int l2_ahead;   // > 0: regions this many iterations ahead are prefetched into L2
  uint64_t goff[1 << (kStageMaxK - 1)];  // request r -> element offset in the state
  uint32_t sreq[1 << (kStageMaxK - 1)];  // request r -> swizzled local offset
  uint32_t smem_j[1 << kStageMaxK];      // member j  -> swizzled local offset
};

B2Q_HD uint32_t stage_swizzle(uint32_t x, const int* src, const int* dst) {
  for (int i = 0; i < 3; ++i) x ^= ((x >> src[i]) & 1u) << dst[i];
  return x;
}

template <int K, bool VEC>
__global__ void __launch_bounds__(kTcThreads, TcTraits<K>::kMinBlocks)
    sv_apply_tc_staged_kernel(const __grid_constant__ TcStagedParams p) {
  constexpr int kTcDim = TcTraits<K>::kDim;
  constexpr int kTcN = TcTraits<K>::kN;
  constexpr int kTcCols = TcTraits<K>::kCols;
  constexpr int kReq = 1 << (K - 1);     // 64-amplitude requests per warp region
  constexpr int kRegion = 1 << (K + 5);  // amplitudes per warp region
  constexpr uint32_t kBufBytes = 4u * kRegion * sizeof(float2);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* sB = reinterpret_cast<float*>(smem_raw);  // B_hi then B_lo
  unsigned char* stage = smem_raw + TcTraits<K>::kBBytes;  // [2 buffers][4 warps][kRegion]
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_s)),
                 "r"((uint32_t)kTcCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    const float4* src = reinterpret_cast<const float4*>(p.bmat);
    float4* dst = reinterpret_cast<float4*>(sB);
    for (int i = tid; i < 2 * kTcN * kTcN / 4; i += kTcThreads) dst[i] = src[i];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
  const uint32_t d_col = 2 * kTcN;
  constexpr uint32_t idesc =
      (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcN >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t sb_hi = smem_u32(sB);
  const uint32_t sb_lo = smem_u32(sB + kTcN * kTcN);
  constexpr uint32_t kLbo = kTcN * 16;
  constexpr uint32_t kSbo = 128;

  // HBM <-> shared: lane l of request r moves local elements (r << 6) + 2l, +1
  const uint64_t lane_goff = (uint64_t)((lane & 15) << 1) + ((uint64_t)(lane >> 4) << p.p5);
  const uint32_t lane_s = stage_swizzle((uint32_t)lane << 1, p.swz_src, p.swz_dst);
  // shared <-> registers: this thread's group
  uint32_t group_local = 0;
#pragma unroll
  for (int i = 0; i < 5; ++i) group_local |= (uint32_t)((lane >> i) & 1) << p.gpos[i];
  const uint32_t sg = stage_swizzle(group_local, p.swz_src, p.swz_dst);
  unsigned char* const stage_warp = stage + (size_t)warp * kRegion * sizeof(float2);
  const uint32_t stage_warp_s = smem_u32(stage_warp);

  auto region_base = [&](uint64_t t) {
    return insert_zero_bits(t * 4 + (uint64_t)warp, p.rpos, K + 5);
  };
  auto prefetch = [&](uint64_t base, uint32_t buf) {
    const float2* src = p.state + base + lane_goff;
    const uint32_t dst = stage_warp_s + buf * kBufBytes;
#pragma unroll
    for (int r = 0; r < kReq; ++r) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + ((lane_s ^ p.sreq[r]) << 3)),
                   "l"(src + p.goff[r])
                   : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  uint32_t parity = 0, buf = 0;
  uint64_t tile = blockIdx.x;
  uint64_t base = 0;
  if (tile < p.num_tiles) {
    base = region_base(tile);
    prefetch(base, 0);
  }
  // lane -> one 256-byte run of the region (L2 prefetch: kReq requests x 2 runs)
  const uint64_t lane_run_off =
      (lane < 2 * kReq) ? p.goff[lane >> 1] + ((uint64_t)(lane & 1) << p.p5) : 0;
  while (tile < p.num_tiles) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    const uint64_t next = tile + gridDim.x;
    uint64_t next_base = 0;
    if (next < p.num_tiles) next_base = region_base(next);
    if (p.early && next < p.num_tiles) prefetch(next_base, buf ^ 1u);
    if (p.l2_ahead > 0) {
      const uint64_t far = tile + (uint64_t)p.l2_ahead * gridDim.x;
      if (far < p.num_tiles && lane < 2 * kReq) {
        const float2* q = p.state + region_base(far) + lane_run_off;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], 256;" ::"l"(q) : "memory");
      }
    }
    unsigned char* const sbuf = stage_warp + (size_t)buf * kBufBytes;
    float2 x[kTcDim];
    if constexpr (VEC) {
#pragma unroll
      for (int i = 0; i < kTcDim / 2; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(sbuf + ((sg ^ p.smem_j[2 * i]) << 3));
        x[2 * i] = make_float2(v.x, v.y);
        x[2 * i + 1] = make_float2(v.z, v.w);
      }
    } else {
#pragma unroll
      for (int j = 0; j < kTcDim; ++j)
        x[j] = *reinterpret_cast<const float2*>(sbuf + ((sg ^ p.smem_j[j]) << 3));
    }
#pragma unroll
    for (int c16 = 0; c16 < kTcN / 16; ++c16) {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int col = c16 * 16 + e;
        const float a = (col & 1) ? x[col >> 1].y : x[col >> 1].x;
        const uint32_t h = to_tf32(a);
        hi[e] = h;
        lo[e] = to_tf32(a - __uint_as_float(h));
      }
      tmem_st16(lane_base + (uint32_t)(c16 * 16), hi);
      tmem_st16(lane_base + (uint32_t)(kTcN + c16 * 16), lo);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    // next region -> the other buffer while the tensor core works on this one
    if (!p.early && next < p.num_tiles) prefetch(next_base, buf ^ 1u);
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // same product order as sv_apply_tc_kernel (see the comment there)
      uint32_t acc0 = 0;
#pragma unroll
      for (int prod = 0; prod < 2; ++prod) {
        const uint32_t a_col = (prod == 0) ? (uint32_t)kTcN : 0u;
        const uint32_t sb = (prod == 0) ? sb_hi : sb_lo;
#pragma unroll
        for (int ks = 0; ks < kTcN / 8; ++ks) {
          const uint64_t bd = umma_desc(sb + (uint32_t)(ks * 2) * kLbo, kLbo, kSbo);
          umma_tf32_ts(tmem_base + d_col, tmem_base + a_col + (uint32_t)(ks * 8), bd, idesc, acc0);
          acc0 = 1;
        }
      }
#pragma unroll
      for (int ks = 0; ks < kTcN / 8; ++ks) {
        const uint64_t bd = umma_desc(sb_hi + (uint32_t)(ks * 2) * kLbo, kLbo, kSbo);
        const bool second = ks >= kTcN / 16;
        umma_tf32_ts(tmem_base + d_col + (second ? (uint32_t)kTcN : 0u),
                     tmem_base + (uint32_t)(ks * 8), bd, idesc,
                     (second && ks == kTcN / 16) ? 0u : 1u);
      }
      asm volatile(
          "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
              smem_u32(&mbar))
          : "memory");
    }
    mbar_wait(smem_u32(&mbar), parity);
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // D0 + D1 -> this thread's slots of the region (all lanes finished reading
    // their inputs before the __syncthreads above)
#pragma unroll
    for (int c16 = 0; c16 < kTcN / 16; ++c16) {
      uint32_t d[16], d1[16];
      tmem_ld16(lane_base + d_col + (uint32_t)(c16 * 16), d);
      tmem_ld16(lane_base + d_col + (uint32_t)(kTcN + c16 * 16), d1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if constexpr (VEC) {
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          const int r = (c16 * 16 + e) >> 1;
          float4 o;
          o.x = __uint_as_float(d[e]) + __uint_as_float(d1[e]);
          o.y = __uint_as_float(d[e + 1]) + __uint_as_float(d1[e + 1]);
          o.z = __uint_as_float(d[e + 2]) + __uint_as_float(d1[e + 2]);
          o.w = __uint_as_float(d[e + 3]) + __uint_as_float(d1[e + 3]);
          *reinterpret_cast<float4*>(sbuf + ((sg ^ p.smem_j[r]) << 3)) = o;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          const int r = (c16 * 16 + e) >> 1;
          float2 o;
          o.x = __uint_as_float(d[e]) + __uint_as_float(d1[e]);
          o.y = __uint_as_float(d[e + 1]) + __uint_as_float(d1[e + 1]);
          *reinterpret_cast<float2*>(sbuf + ((sg ^ p.smem_j[r]) << 3)) = o;
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    // region -> HBM, 512 contiguous bytes per warp request
    {
      constexpr int kBatch = kReq < 8 ? kReq : 8;
      const uint64_t idx0 = base + lane_goff;
#pragma unroll
      for (int r0 = 0; r0 < kReq; r0 += kBatch) {
        float4 v[kBatch];
#pragma unroll
        for (int r = 0; r < kBatch; ++r)
          v[r] = *reinterpret_cast<const float4*>(sbuf + ((lane_s ^ p.sreq[r0 + r]) << 3));
        if (p.xbit < 0) {
#pragma unroll
          for (int r = 0; r < kBatch; ++r)
            *reinterpret_cast<float4*>(p.out_local + idx0 + p.goff[r0 + r]) = v[r];
        } else {
#pragma unroll
          for (int r = 0; r < kBatch; ++r) {
            const uint64_t idx = idx0 + p.goff[r0 + r];
            const bool stays = (int)((idx >> p.xbit) & 1ull) == p.gval;
            const uint64_t idx2 = (idx & ~(1ull << p.xbit)) | ((uint64_t)p.gval << p.xbit);
            float2* const buf = stays ? p.out_local : p.out_peer;
            *reinterpret_cast<float4*>(buf + idx2) = v[r];
          }
        }
      }
    }
    __syncwarp();  // this buffer is refilled by the prefetch issued one iteration later
    tile = next;
    base = next_base;
    buf ^= 1u;
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)kTcCols)
                 : "memory");
  }
}

// Host side of the staged layout (also exported for the host unit tests).
static void make_staged_params(int n, int K, const int* sorted, TcStagedParams* p) {
  bool is_target[64] = {false};
  for (int i = 0; i < K; ++i) is_target[sorted[i]] = true;
  int free_low[5];
  for (int b = 0, c = 0; c < 5; ++b)
    if (!is_target[b]) free_low[c++] = b;
  // region bits, ascending
  int nr = 0;
  {
    bool in_region[64] = {false};
    for (int i = 0; i < K; ++i) in_region[sorted[i]] = true;
    for (int i = 0; i < 5; ++i) in_region[free_low[i]] = true;
    for (int b = 0; b < n; ++b)
      if (in_region[b]) p->rpos[nr++] = b;
  }
  auto local_pos = [&](int bit) {
    for (int i = 0; i < nr; ++i)
      if (p->rpos[i] == bit) return i;
    return -1;
  };
  p->p5 = p->rpos[5];
  for (int i = 0; i < 5; ++i) p->gpos[i] = local_pos(free_low[i]);
  // swizzle: the three lowest group positions >= 1 must land on distinct bits of 1..3
  {
    int cand[3], nc = 0;
    for (int i = 0; i < 5 && nc < 3; ++i)
      if (p->gpos[i] >= 1) cand[nc++] = p->gpos[i];
    bool used[4] = {false, false, false, false};
    for (int i = 0; i < nc; ++i)
      if (cand[i] <= 3) used[cand[i]] = true;
    int ns = 0, next_free = 1;
    for (int i = 0; i < 3; ++i) {
      p->swz_src[i] = 20;
      p->swz_dst[i] = 1;
    }
    for (int i = 0; i < nc; ++i) {
      if (cand[i] <= 3) continue;
      while (next_free <= 3 && used[next_free]) ++next_free;
      p->swz_src[ns] = cand[i];
      p->swz_dst[ns] = next_free;
      used[next_free] = true;
      ++ns;
    }
  }
  const int nreq = 1 << (K - 1);
  for (int r = 0; r < nreq; ++r) {
    uint64_t off = 0;
    for (int i = 0; i < K - 1; ++i)
      if ((r >> i) & 1) off += 1ull << p->rpos[6 + i];
    p->goff[r] = off;
    p->sreq[r] = stage_swizzle((uint32_t)r << 6, p->swz_src, p->swz_dst);
  }
  for (int j = 0; j < (1 << K); ++j) {
    uint32_t loc = 0;
    for (int b = 0; b < K; ++b)
      if ((j >> b) & 1) loc |= 1u << local_pos(sorted[b]);
    p->smem_j[j] = stage_swizzle(loc, p->swz_src, p->swz_dst);
  }
}

// Round-to-nearest-even onto the TF32 grid (10 explicit mantissa bits).
static float tf32_round_host(float x) {
  uint32_t u;
  std::memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return x;
  const uint32_t lsb = (u >> 13) & 1u;
  u += 0xfffu + lsb;
  u &= ~0x1fffu;
  float r;
  std::memcpy(&r, &u, 4);
  return r;
}

std::atomic<int> g_tc_mode{1};  // 0 = off, 1 = tensor cores for k = 5, 6 (complex64), 2 = also k = 4
// 0 = never stage through shared memory, 1 = stage when a target sits on index
// bit 0 or 1, 2 = always (k <= 5)
std::atomic<int> g_tc_stage_mode{1};
std::atomic<int> g_tc_stage_early{0};
std::atomic<int> g_tc_stage_l2_ahead{0};

// Gate matrices travel through a ring of pinned host / device slot pairs: a
// cudaMemcpyAsync from PAGEABLE memory synchronises the stream first, which
// would serialise the host scheduler with the GPU on every pass.
struct MatrixRing {
  static constexpr int kSlots = 64;
  static constexpr size_t kBytes = TcTraits<kTcMaxK>::kBBytes;  // sized for the widest block
  float* host[kSlots] = {nullptr};
  float* dev[kSlots] = {nullptr};
  cudaEvent_t done[kSlots];
  bool used[kSlots] = {false};
  uint64_t next = 0;
  bool ready = false;
};
static MatrixRing g_ring[64];
static std::mutex g_ring_mutex;

static int ring_acquire(MatrixRing** out_ring, int* out_slot) {
  int devid = 0;
  B2Q_CUDA_CHECK(cudaGetDevice(&devid));
  B2Q_REQUIRE(devid >= 0 && devid < 64, "device index out of range");
  MatrixRing& r = g_ring[devid];
  if (!r.ready) {
    for (int i = 0; i < MatrixRing::kSlots; ++i) {
      B2Q_CUDA_CHECK(cudaHostAlloc((void**)&r.host[i], MatrixRing::kBytes, cudaHostAllocDefault));
      B2Q_CUDA_CHECK(cudaMalloc((void**)&r.dev[i], MatrixRing::kBytes));
      B2Q_CUDA_CHECK(cudaEventCreateWithFlags(&r.done[i], cudaEventDisableTiming));
    }
    r.ready = true;
  }
  const int slot = (int)(r.next++ % MatrixRing::kSlots);
  if (r.used[slot]) B2Q_CUDA_CHECK(cudaEventSynchronize(r.done[slot]));
  r.used[slot] = true;
  *out_ring = &r;
  *out_slot = slot;
  return B2Q_OK;
}

// mode 1: k = 5, 6 on the tensor cores; mode 2: also k = 4.
bool tc_applicable(int dtype, int n, int K) {
  const int mode = g_tc_mode.load(std::memory_order_relaxed);
  if (mode == 0 || dtype != B2Q_C64 || n < K + 7) return false;
  return K == 5 || K == 6 || (K == 4 && mode == 2);
}

template <int K, int LOW>
static int launch_tc_kernel(const TcParams& p, cudaStream_t stream) {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  static bool attr_set[64] = {false};
  {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
      B2Q_CUDA_CHECK(cudaFuncSetAttribute(sv_apply_tc_kernel<K, LOW>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)TcTraits<K>::kBBytes));
      attr_set[dev] = true;
    }
  }
  const uint64_t grid =
      std::min<uint64_t>(p.num_tiles, (uint64_t)sms * TcTraits<K>::kMinBlocks);
  sv_apply_tc_kernel<K, LOW><<<(unsigned)grid, kTcThreads, TcTraits<K>::kBBytes, stream>>>(p);
  B2Q_LAUNCH_CHECK("sv_apply_tc_kernel");
  return B2Q_OK;
}

template <int K, bool VEC>
static int launch_tc_staged_kernel(const TcStagedParams& p, cudaStream_t stream) {
  constexpr size_t kSmem = TcTraits<K>::kBBytes + 2ull * 4ull * (1ull << (K + 5)) * sizeof(float2);
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  static bool attr_set[64] = {false};
  {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
      B2Q_CUDA_CHECK(cudaFuncSetAttribute(sv_apply_tc_staged_kernel<K, VEC>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
      attr_set[dev] = true;
    }
  }
  const uint64_t grid =
      std::min<uint64_t>(p.num_tiles, (uint64_t)sms * TcTraits<K>::kMinBlocks);
  sv_apply_tc_staged_kernel<K, VEC><<<(unsigned)grid, kTcThreads, kSmem, stream>>>(p);
  B2Q_LAUNCH_CHECK("sv_apply_tc_staged_kernel");
  return B2Q_OK;
}


// ---- tile kernel: SEVERAL blocks in one pass over HBM -------------------------
//
// A pass costs 2.6 ms of HBM time at 30 qubits however little it computes, and
// the 5-qubit tensor-core pass already runs at ~0.9 of the copy rate — the
// remaining lever is the NUMBER of passes.  Two consecutive 5-qubit blocks
// touch at most 10 index bits; a CTA tile of 2^13 amplitudes (64 KB) over those
// bits plus index bits 0-2 (so the tile moves in runs of >= 64 bytes) and the
// lowest free ones holds every amplitude group of BOTH blocks.  The tile is
// staged once (cp.async, 16-byte lanes), block A and block B are applied to it
// in shared memory — gather a group -> 3xTF32 split -> TMEM -> tcgen05.mma ->
// TMEM -> scatter, exactly the arithmetic of the one-block kernels — and it is
// written back once: half the HBM traffic per block.
//
//   * 8192 amplitudes / 32 per group = 256 groups = TWO UMMA M tiles per block,
//     each with its own half of TMEM (A_hi | A_lo | D0 | D1) and its own
//     mbarrier; they share the tile and the B matrices.  512 threads: TWO per
//     UMMA row (each moves 16 of the group's 32 members, through the two warps
//     that may touch that TMEM lane quarter), so 16 warps hide the shared-memory,
//     conversion and TMEM latencies of one another.  Both M tiles gather at the
//     same time, their MMAs run back to back on the tensor core, and the first
//     one's epilogue overlaps the second one's MMAs.
//   * The tile is double buffered, and all HBM traffic is issued while block A is
//     on the tensor core: the previous tile's write-back and the next tile's copy.  Shared memory: 2 x 32 KB of B + 2 x 64 KB of tile
//     = 192 KB, one CTA per SM, all 512 TMEM columns.
//   * The tile is stored swizzled: slot bits 1-3 (the 8-byte bank pairs) are
//     parities of local-index bits chosen on the host so that the linear copy
//     pattern and the group gather / scatter of BOTH blocks are free of bank
//     conflicts (make_tile_params; replayed on the CPU in tests/test_plan_host.py).
constexpr int kTileBits = 13;
constexpr int kTileAmps = 1 << kTileBits;
constexpr int kTileMaxBlocks = 2;
constexpr int kTileGroups = 2;              // UMMA M tiles per block = TMEM regions
constexpr int kTileGroupThreads = 256;      // two threads per UMMA row: one per half of the members
constexpr int kTileThreads = kTileGroups * kTileGroupThreads;
constexpr int kTileGroupBits = kTileBits - 5;  // bits of the amplitude-group index
constexpr int kTileRounds = kTileAmps / 2 / kTileThreads;  // 16-byte copies per thread and tile
constexpr uint32_t kTileBufBytes = kTileAmps * sizeof(float2);
constexpr size_t kTileBBytes = TcTraits<5>::kBBytes;  // per block: B_hi + B_lo
constexpr size_t kTileSmemBytes = kTileMaxBlocks * kTileBBytes + 2ull * kTileBufBytes;

struct TcTileParams {
  float2* state;
  uint64_t num_tiles;
  const float* bmat;  // num_blocks x (B_hi, B_lo), UMMA K-major layout
  int num_blocks;
  int tbits[kTileBits];  // ascending index positions of the tile bits (tbits[i] = i for i < 3)
  uint32_t xmask[3];     // slot bit d + 1 = parity(local index & xmask[d])
  uint64_t rgoff[kTileRounds];  // copy round r (local bits 10-12) -> element offset in the state
  uint32_t rslot[kTileRounds];  // copy round r -> slot offset
  int vec[kTileMaxBlocks];                        // 1: local bit 0 is a target (members 2i, 2i+1 adjacent)
  int gbit[kTileMaxBlocks][kTileGroupBits];       // local position behind bit i of the group index
  uint32_t mslot[kTileMaxBlocks][32];             // member j -> slot offset
};

B2Q_HD uint32_t tile_slot(uint32_t local, const uint32_t* xmask) {
  uint32_t s = local & ~0xeu;
  for (int d = 0; d < 3; ++d) {
    uint32_t v = local & xmask[d];
    v ^= v >> 16;
    v ^= v >> 8;
    v ^= v >> 4;
    v ^= v >> 2;
    v ^= v >> 1;
    s |= (v & 1u) << (d + 1);
  }
  return s;
}

__device__ __forceinline__ void group_barrier(int group) {
  asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(kTileGroupThreads) : "memory");
}

__global__ void __launch_bounds__(kTileThreads, 1)
    sv_apply_tc_tile_kernel(const __grid_constant__ TcTileParams p) {
  constexpr int kTcDim = 32;
  constexpr int kTcN = 64;
  constexpr int kHalf = kTcDim / 2;     // members per thread
  constexpr int kGroupCols = 4 * kTcN;  // A_hi | A_lo | D0 | D1
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* sB = reinterpret_cast<float*>(smem_raw);
  __shared__ __align__(8) uint64_t mbar[kTileGroups];
  __shared__ uint32_t tmem_base_s;

  // 16 warps: warp = q | h << 2 | g << 3.  q = TMEM lane quarter (the hardware lets a
  // warp touch lanes 32 * (warp % 4) ...), h = which half of a group's 32 members this
  // thread moves, g = UMMA M tile.  Two warps share every lane quarter, so the per-thread
  // gather / convert / epilogue chains are half as long and four warps per scheduler hide
  // each other's latencies.
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int q = warp & 3;
  const int h = (warp >> 2) & 1;
  const int group = warp >> 3;
  const int gidx = lane | (q << 5) | (group << 7);  // amplitude group of this thread

  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_s)),
                 "r"((uint32_t)(kTileGroups * kGroupCols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int g = 0; g < kTileGroups; ++g)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[g])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    const float4* src = reinterpret_cast<const float4*>(p.bmat);
    float4* dst = reinterpret_cast<float4*>(sB);
    const int count = p.num_blocks * (int)(kTileBBytes / sizeof(float4));
    for (int i = tid; i < count; i += kTileThreads) dst[i] = src[i];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_group = tmem_base_s + (uint32_t)(group * kGroupCols);
  const uint32_t lane_base = tmem_group + ((uint32_t)(q * 32) << 16);
  const uint32_t d_col = 2 * kTcN;
  constexpr uint32_t idesc =
      (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcN >> 3) << 17) | ((128u >> 4) << 24);
  constexpr uint32_t kLbo = kTcN * 16;
  constexpr uint32_t kSbo = 128;
  const uint32_t mbar_s = smem_u32(&mbar[group]);
  const bool issuer = (tid & (kTileGroupThreads - 1)) == 0;

  // copy pattern: this thread moves local elements (2 * lane | warp << 6 | round << 10), +1
  const uint32_t thr_local = ((uint32_t)lane << 1) | ((uint32_t)warp << 6);
  uint64_t thr_goff = 0;
#pragma unroll
  for (int i = 1; i < 10; ++i) thr_goff += (uint64_t)((thr_local >> i) & 1u) << p.tbits[i];
  const uint32_t thr_slot = tile_slot(thr_local, p.xmask);
  // gather pattern: this thread's amplitude group, per block
  uint32_t sg[kTileMaxBlocks];
#pragma unroll
  for (int b = 0; b < kTileMaxBlocks; ++b) {
    uint32_t gl = 0;
#pragma unroll
    for (int i = 0; i < kTileGroupBits; ++i) gl |= (uint32_t)((gidx >> i) & 1) << p.gbit[b][i];
    sg[b] = tile_slot(gl, p.xmask);
  }
  unsigned char* const tiles = smem_raw + kTileMaxBlocks * kTileBBytes;
  const uint32_t tiles_s = smem_u32(tiles);

  auto prefetch = [&](uint64_t base, uint32_t buf) {
    const float2* src = p.state + base + thr_goff;
    const uint32_t dst = tiles_s + buf * kTileBufBytes;
#pragma unroll
    for (int r = 0; r < kTileRounds; ++r) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + ((thr_slot ^ p.rslot[r]) << 3)),
                   "l"(src + p.rgoff[r])
                   : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // tile -> HBM, 16-byte lanes, the copy pattern backwards
  auto copy_out = [&](uint64_t tile_base, uint32_t from_buf) {
    float2* const dstp = p.state + tile_base + thr_goff;
    const unsigned char* const src = tiles + (size_t)from_buf * kTileBufBytes;
    float4 v[kTileRounds];
#pragma unroll
    for (int r = 0; r < kTileRounds; ++r)
      v[r] = *reinterpret_cast<const float4*>(src + ((thr_slot ^ p.rslot[r]) << 3));
#pragma unroll
    for (int r = 0; r < kTileRounds; ++r) *reinterpret_cast<float4*>(dstp + p.rgoff[r]) = v[r];
  };

  // The state offset of a tile (13 inserted zero bits: ~150 instructions) is the same for
  // all 512 threads, so ONE lane computes it, two tiles ahead, in the shadow of block A's
  // MMAs, and hands it over through shared memory (every thread computing it at the top
  // of the tile cost ~850 cycles of the critical path, profiles/r2l_tile_trace_*).
  __shared__ uint64_t s_next_base[2];
  uint64_t tile = blockIdx.x;
  uint64_t base = 0, prev_base = 0;
  bool have_prev = false;
  uint32_t parity = 0, buf = 0, it = 0;
  if (tile < p.num_tiles) {
    base = insert_zero_bits(tile, p.tbits, kTileBits);
    prefetch(base, 0);
    if (tid == 0 && tile + gridDim.x < p.num_tiles)
      s_next_base[0] = insert_zero_bits(tile + gridDim.x, p.tbits, kTileBits);
  }
  while (tile < p.num_tiles) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // the tile is complete and visible to the whole CTA
    const uint64_t next = tile + gridDim.x;
    const uint64_t next_base = next < p.num_tiles ? s_next_base[it & 1u] : 0;
    unsigned char* const sbuf = tiles + (size_t)buf * kTileBufBytes;
    for (int b = 0; b < p.num_blocks; ++b) {
      const uint32_t sgb = b == 0 ? sg[0] : sg[1];
      const uint32_t* const mslot = p.mslot[b] + h * kHalf;  // this thread's members
      const bool vec = p.vec[b] != 0;
      float2 x[kHalf];
      if (vec) {
#pragma unroll
        for (int i = 0; i < kHalf / 2; ++i) {
          const float4 v = *reinterpret_cast<const float4*>(sbuf + ((sgb ^ mslot[2 * i]) << 3));
          x[2 * i] = make_float2(v.x, v.y);
          x[2 * i + 1] = make_float2(v.z, v.w);
        }
      } else {
#pragma unroll
        for (int j = 0; j < kHalf; ++j)
          x[j] = *reinterpret_cast<const float2*>(sbuf + ((sgb ^ mslot[j]) << 3));
      }
      // members [16h, 16h + 16) = real columns [32h, 32h + 32) of the A row
#pragma unroll
      for (int c16 = 0; c16 < 2; ++c16) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int col = c16 * 16 + e;
          const float a = (col & 1) ? x[col >> 1].y : x[col >> 1].x;
          const uint32_t hh = to_tf32(a);
          hi[e] = hh;
          lo[e] = to_tf32(a - __uint_as_float(hh));
        }
        tmem_st16(lane_base + (uint32_t)(h * 32 + c16 * 16), hi);
        tmem_st16(lane_base + (uint32_t)(kTcN + h * 32 + c16 * 16), lo);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      group_barrier(group);  // this M tile's A operand is complete
      if (issuer) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // same product order as sv_apply_tc_kernel (see the comment there)
        const uint32_t sb_hi = smem_u32(sB) + (uint32_t)b * (uint32_t)kTileBBytes;
        const uint32_t sb_lo = sb_hi + (uint32_t)(kTcN * kTcN * sizeof(float));
        uint32_t acc0 = 0;
#pragma unroll
        for (int prod = 0; prod < 2; ++prod) {
          const uint32_t a_col = (prod == 0) ? (uint32_t)kTcN : 0u;
          const uint32_t sb = (prod == 0) ? sb_hi : sb_lo;
#pragma unroll
          for (int ks = 0; ks < kTcN / 8; ++ks) {
            const uint64_t bd = umma_desc(sb + (uint32_t)(ks * 2) * kLbo, kLbo, kSbo);
            umma_tf32_ts(tmem_group + d_col, tmem_group + a_col + (uint32_t)(ks * 8), bd, idesc, acc0);
            acc0 = 1;
          }
        }
#pragma unroll
        for (int ks = 0; ks < kTcN / 8; ++ks) {
          const uint64_t bd = umma_desc(sb_hi + (uint32_t)(ks * 2) * kLbo, kLbo, kSbo);
          const bool second = ks >= kTcN / 16;
          umma_tf32_ts(tmem_group + d_col + (second ? (uint32_t)kTcN : 0u),
                       tmem_group + (uint32_t)(ks * 8), bd, idesc,
                       (second && ks == kTcN / 16) ? 0u : 1u);
        }
        asm volatile(
            "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar_s)
            : "memory");
      }
      if (b == 0 && tid == 32 && next + gridDim.x < p.num_tiles)  // (warp 1: not an MMA issuer)
        s_next_base[(it + 1u) & 1u] = insert_zero_bits(next + gridDim.x, p.tbits, kTileBits);
      if (b == 0) {
        // Under block A's MMAs: the PREVIOUS tile (finished, in the other buffer) goes
        // back to HBM and that buffer is refilled with the next tile.  A thread refills
        // exactly the slots it reads (same copy pattern both ways), so no barrier is
        // needed in between: its slots are read into registers first, then the
        // latency-critical loads of the next tile are issued, then the stores.
        float4 v[kTileRounds];
        if (have_prev) {
          const unsigned char* const src = tiles + (size_t)(buf ^ 1u) * kTileBufBytes;
#pragma unroll
          for (int r = 0; r < kTileRounds; ++r)
            v[r] = *reinterpret_cast<const float4*>(src + ((thr_slot ^ p.rslot[r]) << 3));
        }
        if (next < p.num_tiles) prefetch(next_base, buf ^ 1u);
        if (have_prev) {
          float2* const dstp = p.state + prev_base + thr_goff;
#pragma unroll
          for (int r = 0; r < kTileRounds; ++r) *reinterpret_cast<float4*>(dstp + p.rgoff[r]) = v[r];
        }
      }
      mbar_wait(mbar_s, parity);
      parity ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // D0 + D1 (this thread's 32 real columns) -> its 16 slots of the tile; all four
      // TMEM loads are in flight before the one wait
      {
        uint32_t d[2][16], d1[2][16];
#pragma unroll
        for (int c16 = 0; c16 < 2; ++c16) {
          tmem_ld16(lane_base + d_col + (uint32_t)(h * 32 + c16 * 16), d[c16]);
          tmem_ld16(lane_base + d_col + (uint32_t)(kTcN + h * 32 + c16 * 16), d1[c16]);
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c16 = 0; c16 < 2; ++c16) {
          if (vec) {
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
              const int r = (c16 * 16 + e) >> 1;
              float4 o;
              o.x = __uint_as_float(d[c16][e]) + __uint_as_float(d1[c16][e]);
              o.y = __uint_as_float(d[c16][e + 1]) + __uint_as_float(d1[c16][e + 1]);
              o.z = __uint_as_float(d[c16][e + 2]) + __uint_as_float(d1[c16][e + 2]);
              o.w = __uint_as_float(d[c16][e + 3]) + __uint_as_float(d1[c16][e + 3]);
              *reinterpret_cast<float4*>(sbuf + ((sgb ^ mslot[r]) << 3)) = o;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              const int r = (c16 * 16 + e) >> 1;
              float2 o;
              o.x = __uint_as_float(d[c16][e]) + __uint_as_float(d1[c16][e]);
              o.y = __uint_as_float(d[c16][e + 1]) + __uint_as_float(d1[c16][e + 1]);
              *reinterpret_cast<float2*>(sbuf + ((sgb ^ mslot[r]) << 3)) = o;
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();  // both M tiles' results visible before the tile is regrouped
    }
    // (the finished tile stays in its buffer; it is written back under the next
    // tile's first MMAs, or after the loop)
    have_prev = true;
    prev_base = base;
    tile = next;
    base = next_base;
    buf ^= 1u;
    ++it;
  }
  if (have_prev) copy_out(prev_base, buf ^ 1u);
  __syncthreads();
  if (tid < 32) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s),
                 "r"((uint32_t)(kTileGroups * kGroupCols))
                 : "memory");
  }
}

// Host side of the tile layout.  `sorted[b]` = the 5 ascending targets of block b
// (already widened to 5), `tbits` = the 13 ascending tile bits (a superset of all
// targets, containing index bits 0-2).  Returns false if no conflict-free swizzle
// exists (not observed; the caller then applies the blocks one by one).
static bool make_tile_params(int nb, const int (*sorted)[5], const int* tbits, TcTileParams* p) {
  constexpr int G = kTileGroupBits;
  for (int i = 0; i < kTileBits; ++i) p->tbits[i] = tbits[i];
  p->num_blocks = nb;
  auto local_pos = [&](int bit) {
    for (int i = 0; i < kTileBits; ++i)
      if (tbits[i] == bit) return i;
    return -1;
  };
  int tl[kTileMaxBlocks][5];  // local positions of the targets, ascending
  int grp[kTileMaxBlocks][G];  // local positions of the group bits, ascending
  for (int b = 0; b < nb; ++b) {
    bool is_t[kTileBits] = {false};
    for (int i = 0; i < 5; ++i) {
      tl[b][i] = local_pos(sorted[b][i]);
      if (tl[b][i] < 0) return false;
      is_t[tl[b][i]] = true;
    }
    int c = 0;
    for (int l = 0; l < kTileBits; ++l)
      if (!is_t[l]) grp[b][c++] = l;
    p->vec[b] = is_t[0] ? 1 : 0;
  }
  // Bank directions: slot bit d (1..3) of a local index is local bit d XOR the local
  // bits h >= 4 with dir[h] == d.  Every block needs, among its group bits other than
  // local bit 0, one bit per direction (then 16 lanes x 8 bytes, or 8 lanes x 16 bytes,
  // cover all 32 banks).  Small backtracking search over the high bits.
  int dir[kTileBits] = {0};
  struct Need { int b, d; };
  Need needs[3 * kTileMaxBlocks];
  int nn = 0;
  for (int b = 0; b < nb; ++b)
    for (int d = 1; d <= 3; ++d) {
      bool have = false;
      for (int i = 0; i < G; ++i) have |= grp[b][i] == d;
      if (!have) needs[nn++] = Need{b, d};
    }
  int pick[3 * kTileMaxBlocks];
  int depth = 0;
  for (int i = 0; i < nn; ++i) pick[i] = -1;
  bool ok = nn == 0;
  while (nn > 0) {
    if (depth == nn) { ok = true; break; }
    if (depth < 0) break;
    const Need& nd = needs[depth];
    // undo the previous choice at this depth (if it was the one that set dir)
    int start = 0;
    if (pick[depth] >= 0) {
      const int h = grp[nd.b][pick[depth] & 0xff];
      if (pick[depth] & 0x100) dir[h] = 0;
      start = (pick[depth] & 0xff) + 1;
      pick[depth] = -1;
    }
    bool placed = false;
    for (int i = start; i < G; ++i) {
      const int h = grp[nd.b][i];
      if (h < 4) continue;
      if (dir[h] == nd.d) { pick[depth] = i; placed = true; break; }
      if (dir[h] == 0) { dir[h] = nd.d; pick[depth] = i | 0x100; placed = true; break; }
    }
    if (placed) ++depth; else --depth;
  }
  if (!ok) return false;
  for (int d = 1; d <= 3; ++d) {
    uint32_t m = 1u << d;
    for (int h = 4; h < kTileBits; ++h)
      if (dir[h] == d) m |= 1u << h;
    p->xmask[d - 1] = m;
  }
  // thread bits of the gather: [local 0 if it is a group bit], one group bit per
  // direction, then the rest ascending
  for (int b = 0; b < nb; ++b) {
    bool used[kTileBits] = {false};
    int out = 0;
    if (!p->vec[b]) { p->gbit[b][out++] = 0; used[0] = true; }
    for (int d = 1; d <= 3; ++d) {
      int chosen = -1;
      for (int i = 0; i < G && chosen < 0; ++i)
        if (grp[b][i] == d) chosen = d;
      for (int i = 0; i < G && chosen < 0; ++i)
        if (grp[b][i] >= 4 && dir[grp[b][i]] == d && !used[grp[b][i]]) chosen = grp[b][i];
      if (chosen < 0) return false;
      p->gbit[b][out++] = chosen;
      used[chosen] = true;
    }
    for (int i = 0; i < G; ++i)
      if (!used[grp[b][i]]) p->gbit[b][out++] = grp[b][i];
    if (out != G) return false;
    for (int j = 0; j < 32; ++j) {
      uint32_t loc = 0;
      for (int i = 0; i < 5; ++i)
        if ((j >> i) & 1) loc |= 1u << tl[b][i];
      p->mslot[b][j] = tile_slot(loc, p->xmask);
    }
  }
  for (int b = nb; b < kTileMaxBlocks; ++b) {
    p->vec[b] = 0;
    for (int i = 0; i < G; ++i) p->gbit[b][i] = i;
    for (int j = 0; j < 32; ++j) p->mslot[b][j] = 0;
  }
  for (int r = 0; r < kTileRounds; ++r) {
    uint64_t off = 0;
    for (int i = 0; i < 3; ++i)
      if ((r >> i) & 1) off += 1ull << tbits[10 + i];
    p->rgoff[r] = off;
    p->rslot[r] = tile_slot((uint32_t)r << 10, p->xmask);
  }
  return true;
}

// Tile bits of a group of blocks: the union of their targets, index bits 0-2,
// and the lowest other bits up to 13.  Returns false if the union is too wide.
bool tile_bits_for(int n, int nb, const int* ks, const int* targets, int* tbits) {
  if (n < kTileBits || nb < 1 || nb > kTileMaxBlocks) return false;
  bool in[64] = {false};
  // index bits 0-2 are always tile bits: the tile then moves in runs of >= 64
  // bytes (with 32-byte runs a pass fell to 1.1 TB/s, profiles/README.md r2b)
  in[0] = in[1] = in[2] = true;
  int count = 3;
  size_t off = 0;
  for (int b = 0; b < nb; ++b) {
    if (ks[b] < 1 || ks[b] > 5) return false;
    for (int i = 0; i < ks[b]; ++i) {
      const int t = targets[off + i];
      if (t < 0 || t >= n) return false;
      if (!in[t]) { in[t] = true; ++count; }
    }
    off += ks[b];
  }
  if (count > kTileBits) return false;
  for (int bit = 0; bit < n && count < kTileBits; ++bit)
    if (!in[bit]) { in[bit] = true; ++count; }
  int c = 0;
  for (int bit = 0; bit < n; ++bit)
    if (in[bit]) tbits[c++] = bit;
  return c == kTileBits;
}

static void fill_b_matrix(const float* mat, float* hi, float* lo) {
  constexpr int kTcDim = 32, kTcN = 64;
  for (int r = 0; r < kTcDim; ++r)
    for (int c = 0; c < kTcDim; ++c) {
      const float mr = mat[2 * (r * kTcDim + c)], mi = mat[2 * (r * kTcDim + c) + 1];
      const float vals[2][2] = {{mr, -mi}, {mi, mr}};
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
          const int nn = 2 * r + a, kk = 2 * c + b;
          const size_t idx = (size_t)(kk >> 2) * (kTcN * 4) + (size_t)nn * 4 + (kk & 3);
          const float v = vals[a][b];
          const float h = tf32_round_host(v);
          hi[idx] = h;
          lo[idx] = tf32_round_host(v - h);
        }
    }
}

// `mats[b]` = block b's matrix in sorted-target order (plain (re, im) pairs, 32 x 32).
int launch_tc_tile(void* state, int n, int nb, const int (*sorted)[5], const int* tbits,
                   const float* const* mats, cudaStream_t stream) {
  TcTileParams p;
  memset(&p, 0, sizeof(p));
  if (!make_tile_params(nb, sorted, tbits, &p))
    return set_error(B2Q_ERR_UNSUPPORTED, "no conflict-free tile layout for these blocks");
  std::lock_guard<std::mutex> lock(g_ring_mutex);
  MatrixRing* ring = nullptr;
  int slot = 0;
  {
    const int rc = ring_acquire(&ring, &slot);
    if (rc != B2Q_OK) return rc;
  }
  static_assert(kTileMaxBlocks * kTileBBytes <= MatrixRing::kBytes, "ring slot too small");
  for (int b = 0; b < nb; ++b) {
    float* hi = ring->host[slot] + (size_t)b * (kTileBBytes / sizeof(float));
    fill_b_matrix(mats[b], hi, hi + 64 * 64);
  }
  B2Q_CUDA_CHECK(cudaMemcpyAsync(ring->dev[slot], ring->host[slot], nb * kTileBBytes,
                                 cudaMemcpyHostToDevice, stream));
  p.state = reinterpret_cast<float2*>(state);
  p.num_tiles = 1ull << (n - kTileBits);
  p.bmat = ring->dev[slot];
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  static bool attr_set[64] = {false};
  {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
      B2Q_CUDA_CHECK(cudaFuncSetAttribute(sv_apply_tc_tile_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)kTileSmemBytes));
      attr_set[dev] = true;
    }
  }
  const uint64_t grid = std::min<uint64_t>(p.num_tiles, (uint64_t)sms);
  sv_apply_tc_tile_kernel<<<(unsigned)grid, kTileThreads, kTileSmemBytes, stream>>>(p);
  B2Q_LAUNCH_CHECK("sv_apply_tc_tile_kernel");
  B2Q_CUDA_CHECK(cudaEventRecord(ring->done[slot], stream));
  return B2Q_OK;
}

struct TcExchange {
  void* out_local;
  void* out_peer;
  int xbit;
  int gval;
};

// `mat` = gate matrix in sorted-target order (index bit i <-> i-th lowest
// target), plain (re, im) float pairs, row-major 2^K x 2^K.  `ex` != nullptr:
// out-of-place pass fused with a qubit exchange (staged kernel only).
template <int K>
int launch_tc_k(void* state, int n, const int* sorted, const float* mat, cudaStream_t stream,
                const TcExchange* ex = nullptr) {
  constexpr int kTcK = K;
  constexpr int kTcDim = TcTraits<K>::kDim;
  constexpr int kTcN = TcTraits<K>::kN;
  // B[nn][kk], nn = 2r + {0: re, 1: im} of output row r, kk = 2c + {0: re, 1: im}
  // of input column c:  out_re = Mr x_re - Mi x_im ; out_im = Mi x_re + Mr x_im
  std::lock_guard<std::mutex> lock(g_ring_mutex);
  MatrixRing* ring = nullptr;
  int slot = 0;
  {
    const int rc = ring_acquire(&ring, &slot);
    if (rc != B2Q_OK) return rc;
  }
  float* hi = ring->host[slot];
  float* lo = hi + kTcN * kTcN;
  for (int r = 0; r < kTcDim; ++r)
    for (int c = 0; c < kTcDim; ++c) {
      const float mr = mat[2 * (r * kTcDim + c)], mi = mat[2 * (r * kTcDim + c) + 1];
      const float vals[2][2] = {{mr, -mi}, {mi, mr}};
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
          const int nn = 2 * r + a, kk = 2 * c + b;
          const size_t idx = (size_t)(kk >> 2) * (kTcN * 4) + (size_t)nn * 4 + (kk & 3);
          const float v = vals[a][b];
          const float h = tf32_round_host(v);
          hi[idx] = h;
          lo[idx] = tf32_round_host(v - h);
        }
    }
  float* dmat = ring->dev[slot];
  B2Q_CUDA_CHECK(cudaMemcpyAsync(dmat, ring->host[slot], TcTraits<K>::kBBytes,
                                 cudaMemcpyHostToDevice, stream));
  TcParams p;
  p.state = reinterpret_cast<float2*>(state);
  p.num_tiles = (1ull << (n - kTcK)) / kTcThreads;
  for (int i = 0; i < kTcK; ++i) p.tpos[i] = sorted[i];
  p.bmat = dmat;
  const int low = sorted[0] == 0 ? 1 : (sorted[0] == 1 ? 2 : 0);
  const int stage_mode = g_tc_stage_mode.load(std::memory_order_relaxed);
  int rc;
  if constexpr (K <= kStageMaxK) {
    if (ex != nullptr || stage_mode == 2 || (stage_mode == 1 && low != 0)) {
      TcStagedParams sp;
      sp.state = p.state;
      sp.num_tiles = p.num_tiles;
      sp.bmat = dmat;
      make_staged_params(n, K, sorted, &sp);
      sp.out_local = ex ? reinterpret_cast<float2*>(ex->out_local) : p.state;
      sp.out_peer = ex ? reinterpret_cast<float2*>(ex->out_peer) : nullptr;
      sp.xbit = ex ? ex->xbit : -1;
      sp.gval = ex ? ex->gval : 0;
      sp.early = g_tc_stage_early.load(std::memory_order_relaxed);
      sp.l2_ahead = g_tc_stage_l2_ahead.load(std::memory_order_relaxed);
      rc = low == 1 ? launch_tc_staged_kernel<K, true>(sp, stream)
                    : launch_tc_staged_kernel<K, false>(sp, stream);
      if (rc != B2Q_OK) return rc;
      B2Q_CUDA_CHECK(cudaEventRecord(ring->done[slot], stream));
      return B2Q_OK;
    }
  }
  if (ex != nullptr) return set_error(B2Q_ERR_INVALID, "exchange needs the staged kernel (k <= 5)");
  if (low == 1)
    rc = launch_tc_kernel<K, 1>(p, stream);
  else if (low == 2)
    rc = launch_tc_kernel<K, 2>(p, stream);
  else
    rc = launch_tc_kernel<K, 0>(p, stream);
  if (rc != B2Q_OK) return rc;
  B2Q_CUDA_CHECK(cudaEventRecord(ring->done[slot], stream));
  return B2Q_OK;
}

int launch_tc_exchange(void* state, int n, int K, const int* sorted, const float* mat,
                       void* out_local, void* out_peer, int xbit, int gval, cudaStream_t stream) {
  const TcExchange ex{out_local, out_peer, xbit, gval};
  if (K == 4) return launch_tc_k<4>(state, n, sorted, mat, stream, &ex);
  return launch_tc_k<5>(state, n, sorted, mat, stream, &ex);
}

int launch_tc(void* state, int n, int K, const int* sorted, const float* mat, cudaStream_t stream) {
  if (K == 4) return launch_tc_k<4>(state, n, sorted, mat, stream);
  if (K == 5) return launch_tc_k<5>(state, n, sorted, mat, stream);
  return launch_tc_k<6>(state, n, sorted, mat, stream);
}

}  // namespace b2q

extern "C" int b2q_set_tc_stage_opts(int early, int l2_ahead) {
  B2Q_REQUIRE(early >= 0 && early <= 1 && l2_ahead >= 0 && l2_ahead <= 8, "bad stage options");
  b2q::g_tc_stage_early.store(early, std::memory_order_relaxed);
  b2q::g_tc_stage_l2_ahead.store(l2_ahead, std::memory_order_relaxed);
  return B2Q_OK;
}

extern "C" int b2q_set_tc_stage_mode(int mode) {
  B2Q_REQUIRE(mode >= 0 && mode <= 2, "tc stage mode must be 0, 1 or 2");
  b2q::g_tc_stage_mode.store(mode, std::memory_order_relaxed);
  return B2Q_OK;
}

// out: rpos[10] | p5 | gpos[5] | swz_src[3] | swz_dst[3] | goff[16] | sreq[16] | smem_j[32]
extern "C" int b2q_debug_tc_stage_plan(int n_qubits, const int* sorted_targets, int k, int64_t* out) {
  B2Q_REQUIRE(sorted_targets != nullptr && out != nullptr, "null argument");
  B2Q_REQUIRE(k >= 2 && k <= b2q::kStageMaxK && n_qubits >= k + 7, "unsupported shape");
  for (int i = 0; i < k; ++i)
    B2Q_REQUIRE(sorted_targets[i] >= 0 && sorted_targets[i] < n_qubits &&
                    (i == 0 || sorted_targets[i] > sorted_targets[i - 1]),
                "targets must be ascending and in range");
  b2q::TcStagedParams p;
  memset(&p, 0, sizeof(p));
  b2q::make_staged_params(n_qubits, k, sorted_targets, &p);
  int o = 0;
  for (int i = 0; i < 10; ++i) out[o++] = i < k + 5 ? p.rpos[i] : -1;
  out[o++] = p.p5;
  for (int i = 0; i < 5; ++i) out[o++] = p.gpos[i];
  for (int i = 0; i < 3; ++i) out[o++] = p.swz_src[i];
  for (int i = 0; i < 3; ++i) out[o++] = p.swz_dst[i];
  for (int i = 0; i < 16; ++i) out[o++] = i < (1 << (k - 1)) ? (int64_t)p.goff[i] : -1;
  for (int i = 0; i < 16; ++i) out[o++] = i < (1 << (k - 1)) ? (int64_t)p.sreq[i] : -1;
  for (int i = 0; i < 32; ++i) out[o++] = i < (1 << k) ? (int64_t)p.smem_j[i] : -1;
  return B2Q_OK;
}

// Host-only: the tile kernel's address tables for `num_blocks` blocks of 5 ascending
// targets each.  out (int64): tbits[13] | xmask[3] | rgoff[8] | rslot[8] | then per
// block: vec | gbit[8] | mslot[32].  Returns B2Q_ERR_UNSUPPORTED if the blocks do not
// fit one tile.
extern "C" int b2q_debug_tile_plan(int n_qubits, int num_blocks, const int* sorted_targets,
                                   int64_t* out) {
  B2Q_REQUIRE(sorted_targets != nullptr && out != nullptr, "null argument");
  B2Q_REQUIRE(num_blocks >= 1 && num_blocks <= b2q::kTileMaxBlocks, "bad block count");
  int ks[b2q::kTileMaxBlocks];
  for (int b = 0; b < num_blocks; ++b) ks[b] = 5;
  int tbits[b2q::kTileBits];
  if (!b2q::tile_bits_for(n_qubits, num_blocks, ks, sorted_targets, tbits))
    return b2q::set_error(B2Q_ERR_UNSUPPORTED, "blocks do not fit one tile");
  int sorted[b2q::kTileMaxBlocks][5];
  for (int b = 0; b < num_blocks; ++b)
    for (int i = 0; i < 5; ++i) sorted[b][i] = sorted_targets[5 * b + i];
  b2q::TcTileParams p;
  memset(&p, 0, sizeof(p));
  if (!b2q::make_tile_params(num_blocks, sorted, tbits, &p))
    return b2q::set_error(B2Q_ERR_UNSUPPORTED, "no conflict-free tile layout");
  int o = 0;
  for (int i = 0; i < b2q::kTileBits; ++i) out[o++] = p.tbits[i];
  for (int i = 0; i < 3; ++i) out[o++] = p.xmask[i];
  for (int i = 0; i < b2q::kTileRounds; ++i) out[o++] = (int64_t)p.rgoff[i];
  for (int i = 0; i < b2q::kTileRounds; ++i) out[o++] = p.rslot[i];
  for (int b = 0; b < num_blocks; ++b) {
    out[o++] = p.vec[b];
    for (int i = 0; i < b2q::kTileGroupBits; ++i) out[o++] = p.gbit[b][i];
    for (int j = 0; j < 32; ++j) out[o++] = p.mslot[b][j];
  }
  return B2Q_OK;
}

extern "C" int b2q_set_tc_mode(int mode) {
  B2Q_REQUIRE(mode >= 0 && mode <= 2, "tc mode must be 0, 1 or 2");
  b2q::g_tc_mode.store(mode, std::memory_order_relaxed);
  return B2Q_OK;
}
