// Library-level entry points of the cirq_b200 C-ABI (include/cirq_b200.h).
#include "b2q_common.cuh"

#include <mutex>

namespace b2q {

std::atomic<uint64_t> g_launch_count{0};

char* last_error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

namespace {
struct DeviceScratch {
  void* ptr = nullptr;
  size_t bytes = 0;
};
std::mutex g_scratch_mutex;
DeviceScratch g_scratch[64];
}  // namespace

void* workspace(size_t bytes) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    set_error(B2Q_ERR_CUDA, "cudaGetDevice failed");
    return nullptr;
  }
  std::lock_guard<std::mutex> lock(g_scratch_mutex);
  DeviceScratch& s = g_scratch[dev];
  if (s.bytes >= bytes) return s.ptr;
  size_t want = bytes < (1u << 20) ? (1u << 20) : bytes;
  if (s.ptr != nullptr) {
    cudaDeviceSynchronize();
    cudaFree(s.ptr);
    s.ptr = nullptr;
    s.bytes = 0;
  }
  if (cudaMalloc(&s.ptr, want) != cudaSuccess) {
    set_error(B2Q_ERR_CUDA, "cudaMalloc of %zu scratch bytes failed", want);
    s.ptr = nullptr;
    return nullptr;
  }
  s.bytes = want;
  return s.ptr;
}

}  // namespace b2q

extern "C" int b2q_version(void) { return 1000; }

extern "C" const char* b2q_last_error(void) { return b2q::last_error_buffer(); }

extern "C" uint64_t b2q_launch_count(void) {
  return b2q::g_launch_count.load(std::memory_order_relaxed);
}

extern "C" int b2q_device_info(int* sm_count, uint64_t* hbm_bytes, int* cc) {
  int dev = 0;
  B2Q_CUDA_CHECK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  B2Q_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (hbm_bytes) *hbm_bytes = (uint64_t)prop.totalGlobalMem;
  if (cc) *cc = prop.major * 10 + prop.minor;
  return B2Q_OK;
}
