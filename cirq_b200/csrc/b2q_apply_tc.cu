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

template <int K>
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
  auto issue_loads = [&](uint64_t t) {
    const uint64_t g = t * kTcThreads + (uint64_t)tid;
    ptr = p.state + insert_zero_bits(g, p.tpos, kTcK);
#pragma unroll
    for (int j = 0; j < kTcDim; ++j) {
      const float2* q = ptr + s_off[j];
      asm volatile("ld.global.v2.f32 {%0,%1}, [%2];" : "=f"(x[j].x), "=f"(x[j].y) : "l"(q));
    }
  };
  if (tile < p.num_tiles) issue_loads(tile);
  while (tile < p.num_tiles) {
    float2* const cur = ptr;
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
#pragma unroll
      for (int e = 0; e < 16; e += 2) {
        const int r = (c16 * 16 + e) >> 1;
        float2* q = cur + s_off[r];
        const float re = __uint_as_float(d[e]) + __uint_as_float(d1[e]);
        const float im = __uint_as_float(d[e + 1]) + __uint_as_float(d1[e + 1]);
        asm volatile("st.global.v2.f32 [%0], {%1,%2};" ::"l"(q), "f"(re), "f"(im) : "memory");
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

std::atomic<int> g_tc_mode{1};  // 1 = use the tensor-core kernel for k = 5 (complex64)

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

// `mat` = gate matrix in sorted-target order (index bit i <-> i-th lowest
// target), plain (re, im) float pairs, row-major 2^K x 2^K.
template <int K>
int launch_tc_k(void* state, int n, const int* sorted, const float* mat, cudaStream_t stream) {
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
      B2Q_CUDA_CHECK(cudaFuncSetAttribute(sv_apply_tc_kernel<K>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)TcTraits<K>::kBBytes));
      attr_set[dev] = true;
    }
  }
  const uint64_t grid =
      std::min<uint64_t>(p.num_tiles, (uint64_t)sms * TcTraits<K>::kMinBlocks);
  sv_apply_tc_kernel<K><<<(unsigned)grid, kTcThreads, TcTraits<K>::kBBytes, stream>>>(p);
  B2Q_LAUNCH_CHECK("sv_apply_tc_kernel");
  B2Q_CUDA_CHECK(cudaEventRecord(ring->done[slot], stream));
  return B2Q_OK;
}

int launch_tc(void* state, int n, int K, const int* sorted, const float* mat, cudaStream_t stream) {
  if (K == 4) return launch_tc_k<4>(state, n, sorted, mat, stream);
  if (K == 5) return launch_tc_k<5>(state, n, sorted, mat, stream);
  return launch_tc_k<6>(state, n, sorted, mat, stream);
}

}  // namespace b2q

extern "C" int b2q_set_tc_mode(int mode) {
  B2Q_REQUIRE(mode >= 0 && mode <= 2, "tc mode must be 0, 1 or 2");
  b2q::g_tc_mode.store(mode, std::memory_order_relaxed);
  return B2Q_OK;
}
