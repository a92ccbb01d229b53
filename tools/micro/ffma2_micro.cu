// Micro-benchmark: fp32 FMA throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
// Evidence for DESIGN.md section 5 (why the FFMA-bound conv kernels use the packed form).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/ffma2_micro tools/micro/ffma2_micro.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int CHAINS = 16;   // independent accumulators (pairs) per thread
constexpr int ITERS = 4096;

__global__ void __launch_bounds__(256) k_ffma(float* out, float a, float b) {
  float acc[2 * CHAINS];
#pragma unroll
  for (int i = 0; i < 2 * CHAINS; ++i) acc[i] = (float)(threadIdx.x + i);
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 2 * CHAINS; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 2 * CHAINS; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_ffma2(float* out, float a, float b) {
  uint64_t acc[CHAINS];
  uint64_t pa, pb;
  asm("mov.b64 %0, {%1, %1};" : "=l"(pa) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(pb) : "f"(b));
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) {
    const float lo = (float)(threadIdx.x + 2 * i), hi = (float)(threadIdx.x + 2 * i + 1);
    asm("mov.b64 %0, {%1, %2};" : "=l"(acc[i]) : "f"(lo), "f"(hi));
  }
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i)
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(pa), "l"(pb));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
    s += lo + hi;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  const int blocks = sms * 4, threads = 256;
  float* out;
  cudaMalloc(&out, sizeof(float) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  for (int variant = 0; variant < 2; ++variant) {
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      if (variant == 0) k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f);
      else k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    const double fmas = (double)blocks * threads * 2 * CHAINS * ITERS;
    const double tflops = 2.0 * fmas / (best * 1e-3) / 1e12;
    const double per_clk_sm = fmas / (best * 1e-3) / ((double)clk_khz * 1e3) / sms;
    printf("{\"variant\": \"%s\", \"ms\": %.4f, \"tflops\": %.2f, \"fma_per_clk_per_sm_at_max_clock\": %.1f, \"sms\": %d, \"max_clock_mhz\": %d}\n",
           variant == 0 ? "FFMA" : "FFMA2", best, tflops, per_clk_sm, sms, clk_khz / 1000);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
