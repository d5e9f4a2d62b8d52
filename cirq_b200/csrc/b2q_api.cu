// Library-level entry points of the cirq_b200 C-ABI (include/cirq_b200.h).
#include "b2q_common.cuh"

namespace b2q {

std::atomic<uint64_t> g_launch_count{0};

char* last_error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
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
