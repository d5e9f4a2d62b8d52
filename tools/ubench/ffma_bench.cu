// Measures FFMA vs FFMA2 (fma.rn.f32x2) issue throughput per SM on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma_bench ffma_bench.cu
#include <cuda_runtime.h>
#include <cstdio>

template <int CHAINS>
__global__ void k_ffma(float* out, float a, float b, int iters) {
  float acc[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
__global__ void k_ffma2(float2* out, float2 a, float2 b, int iters) {
  float2 acc[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc[i] = __ffma2_rn(acc[i], a, b);
  }
  float2 s = make_float2(0, 0);
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) { s.x += acc[i].x; s.y += acc[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t s, e;
  cudaEventCreate(&s); cudaEventCreate(&e);
  f();
  cudaEventRecord(s);
  f();
  cudaEventRecord(e);
  cudaEventSynchronize(e);
  float ms; cudaEventElapsedTime(&ms, s, e);
  return ms;
}

int main() {
  int sms = 148, threads = 1024, iters = 20000;
  float* out; cudaMalloc(&out, sizeof(float2) * sms * threads * 2);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); sms = p.multiProcessorCount;
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("SMs %d, max clock %.0f MHz\n", sms, clk_khz / 1e3);
  for (int occ = 1; occ <= 2; ++occ) {
    int blocks = sms * occ;
    float ms1 = time_ms([&] { k_ffma<8><<<blocks, threads>>>(out, 1.0001f, 1e-7f, iters); });
    double fma1 = (double)blocks * threads * 8 * iters;
    printf("FFMA  chains=8 blocks/SM=%d: %.3f ms  %.1f GFMA/s  = %.1f lane-FMA/clk/SM @1.9GHz\n", occ, ms1,
           fma1 / ms1 / 1e6, fma1 / ms1 / 1e6 / sms / 1.9);
    float ms2 = time_ms([&] { k_ffma2<8><<<blocks, threads>>>((float2*)out, make_float2(1.0001f, 0.9999f), make_float2(1e-7f, 1e-7f), iters); });
    double fma2 = (double)blocks * threads * 8 * iters * 2;
    printf("FFMA2 chains=8 blocks/SM=%d: %.3f ms  %.1f GFMA/s  = %.1f lane-FMA/clk/SM @1.9GHz\n", occ, ms2,
           fma2 / ms2 / 1e6, fma2 / ms2 / 1e6 / sms / 1.9);
  }
  float ms3 = time_ms([&] { k_ffma2<2><<<sms, 256>>>((float2*)out, make_float2(1.0001f, 0.9999f), make_float2(1e-7f, 1e-7f), iters); });
  printf("FFMA2 chains=2, 8 warps/SM (latency bound): %.3f ms -> %.2f cycles per dependent FFMA2 @1.9GHz\n", ms3,
         ms3 * 1e-3 * 1.9e9 / iters / 2);
  return 0;
}
