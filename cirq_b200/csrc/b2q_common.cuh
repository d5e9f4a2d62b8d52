// Shared helpers for the cirq_b200 CUDA library: error plumbing, bit-index
// arithmetic (host+device so the host unit tests exercise the exact code the
// kernels run), complex arithmetic on float2/double2.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/cirq_b200.h"

#define B2Q_HD __host__ __device__ __forceinline__

namespace b2q {

// ---- error plumbing ---------------------------------------------------------

char* last_error_buffer();  // thread-local, 512 bytes
// Persistent per-device scratch of at least `bytes` (partials, scalars, small
// tables).  Grown on demand; contents are only valid within one library call.
// Returns nullptr (and sets the error) on allocation failure.
void* workspace(size_t bytes);
extern std::atomic<uint64_t> g_launch_count;

inline int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

#define B2Q_CUDA_CHECK(expr)                                                              \
  do {                                                                                    \
    cudaError_t err__ = (expr);                                                           \
    if (err__ != cudaSuccess) {                                                           \
      return b2q::set_error(B2Q_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                 \
                            cudaGetErrorString(err__), __FILE__, __LINE__);               \
    }                                                                                     \
  } while (0)

// Small stream-ordered temporaries (gate tables uploaded next to the kernel that reads
// them) come from the device's default memory pool.  By default the pool hands its memory
// back to the driver at every synchronisation, so each later cudaMallocAsync maps fresh
// pages — cheap on an empty GPU, but measured at up to 0.6 s per call next to a 128 GiB
// state (profiles/README.md r2z).  Keep what the pool has.
inline void keep_async_pool_memory() {
  static bool done[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  done[dev] = true;
}

#define B2Q_LAUNCH_CHECK(name)                                                            \
  do {                                                                                    \
    b2q::g_launch_count.fetch_add(1, std::memory_order_relaxed);                          \
    cudaError_t err__ = cudaGetLastError();                                               \
    if (err__ != cudaSuccess) {                                                           \
      return b2q::set_error(B2Q_ERR_CUDA, "launch of %s failed: %s (%s:%d)", name,        \
                            cudaGetErrorString(err__), __FILE__, __LINE__);               \
    }                                                                                     \
  } while (0)

#define B2Q_REQUIRE(cond, ...)                                                            \
  do {                                                                                    \
    if (!(cond)) return b2q::set_error(B2Q_ERR_INVALID, __VA_ARGS__);                     \
  } while (0)

// ---- bit-index arithmetic ---------------------------------------------------

// Inserts a zero bit at position `pos` of `x` (bits >= pos move up by one).
B2Q_HD uint64_t insert_zero_bit(uint64_t x, int pos) {
  const uint64_t low = x & ((1ull << pos) - 1ull);
  return ((x >> pos) << (pos + 1)) | low;
}

// Inserts zero bits at each of the `count` ASCENDING positions `pos[]`.
B2Q_HD uint64_t insert_zero_bits(uint64_t x, const int* pos, int count) {
  for (int i = 0; i < count; ++i) x = insert_zero_bit(x, pos[i]);
  return x;
}

// Gathers the bits of `x` at `bits[0..m)` into an m-bit value, bits[0] = MSB.
B2Q_HD uint64_t extract_bits_msb_first(uint64_t x, const int* bits, int m) {
  uint64_t v = 0;
  for (int q = 0; q < m; ++q) v = (v << 1) | ((x >> bits[q]) & 1ull);
  return v;
}

// ---- complex helpers --------------------------------------------------------

template <typename real>
struct Cplx;
template <>
struct Cplx<float> {
  using type = float2;
};
template <>
struct Cplx<double> {
  using type = double2;
};

template <typename real>
B2Q_HD typename Cplx<real>::type make_c(real re, real im) {
  typename Cplx<real>::type r;
  r.x = re;
  r.y = im;
  return r;
}

// acc += m * x (complex), m given as (mr, mi)
template <typename real, typename C>
B2Q_HD void cmac(C& acc, real mr, real mi, const C& x) {
  acc.x = fma(mr, x.x, acc.x);
  acc.x = fma(-mi, x.y, acc.x);
  acc.y = fma(mr, x.y, acc.y);
  acc.y = fma(mi, x.x, acc.y);
}

template <typename real, typename C>
B2Q_HD C cmul(real mr, real mi, const C& x) {
  C r;
  r.x = mr * x.x - mi * x.y;
  r.y = mr * x.y + mi * x.x;
  return r;
}

inline size_t elem_bytes(int dtype) { return dtype == B2Q_C64 ? 8 : 16; }

}  // namespace b2q
