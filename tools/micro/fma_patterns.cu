// Micro-benchmark of the register-operand patterns of the conv inner loops (no memory traffic):
// which FMA form sustains what fraction of the fp32 peak on sm_100a when every instruction
// reads fresh accumulators and operands, as the real kernels do.
//   A  : wgrad pattern   - 3 x outer4  (24 FFMA2 / step : scalar x, dy pair reused 12x)
//   As : same with scalar FFMA (48 / step)
//   B  : fwd pattern     - 12 x dot4x4 (96 FFMA2 / step : scalar x, weight pair, 4-deep chains)
//   Bs : same with scalar FFMA
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I nas_3d_unet_b200/csrc -o tools/micro/fma_patterns tools/micro/fma_patterns.cu
#include <cstdio>
#include <cuda_runtime.h>

template <bool PACKED>
struct F {
  static __device__ __forceinline__ void fma2(float2& d, float s, float wx, float wy) {
    if (PACKED) {
      asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %2};\n\tmov.b64 rb, {%3, %4};\n\t"
          "mov.b64 rd, {%0, %1};\n\tfma.rn.f32x2 rd, ra, rb, rd;\n\tmov.b64 {%0, %1}, rd;\n\t}\n"
          : "+f"(d.x), "+f"(d.y) : "f"(s), "f"(wx), "f"(wy));
    } else {
      d.x = fmaf(s, wx, d.x);
      d.y = fmaf(s, wy, d.y);
    }
  }
  static __device__ __forceinline__ void axpy4(float2 (&a)[2], float s, const float4& w) {
    fma2(a[0], s, w.x, w.y);
    fma2(a[1], s, w.z, w.w);
  }
  static __device__ __forceinline__ void outer4(float2 (&a)[4][2], const float4& x, const float4& w) {
    axpy4(a[0], x.x, w); axpy4(a[1], x.y, w); axpy4(a[2], x.z, w); axpy4(a[3], x.w, w);
  }
  static __device__ __forceinline__ void dot4x4(float2 (&a)[2], const float4& x, const float4 (&w)[4]) {
    axpy4(a, x.x, w[0]); axpy4(a, x.y, w[1]); axpy4(a, x.z, w[2]); axpy4(a, x.w, w[3]);
  }
};

constexpr int ITERS = 2048;

template <bool PACKED>
__global__ void __launch_bounds__(256) k_wgrad_like(float* out, const float4* in) {
  float2 acc[3][4][2];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = make_float2(0.f, 0.f);
  float4 x0 = in[threadIdx.x], x1 = in[threadIdx.x + 256], x2 = in[threadIdx.x + 512], g = in[threadIdx.x + 768];
#pragma unroll 1
  for (int it = 0; it < ITERS; it += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      F<PACKED>::outer4(acc[0], x0, g);
      F<PACKED>::outer4(acc[1], x1, g);
      F<PACKED>::outer4(acc[2], x2, g);
      const float4 t = x0; x0 = x1; x1 = x2; x2 = g; g = t;   // rotate (pure renaming when unrolled x4)
    }
  }
  float s = 0.f;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) s += acc[a][b][0].x + acc[a][b][0].y + acc[a][b][1].x + acc[a][b][1].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <bool PACKED>
__global__ void __launch_bounds__(256) k_fwd_like(float* out, const float4* in) {
  float2 acc[4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a) acc[a][0] = acc[a][1] = make_float2(0.f, 0.f);
  float4 w[3][4];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) w[a][b] = in[threadIdx.x + 32 * (a * 4 + b)];
  float4 x[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) x[r] = in[threadIdx.x + 700 + 32 * r];
#pragma unroll 1
  for (int it = 0; it < ITERS; it += 6) {
#pragma unroll
    for (int u = 0; u < 6; ++u) {
#pragma unroll
      for (int oh = 0; oh < 4; ++oh)
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) F<PACKED>::dot4x4(acc[oh], x[(oh + kh + u) % 6], w[kh]);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a) s += acc[a][0].x + acc[a][0].y + acc[a][1].x + acc[a][1].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K>
static void run(const char* name, K kern, double fma_per_thread, int ctas_per_sm, float* out, const float4* in) {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int blocks = p.multiProcessorCount * ctas_per_sm;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    kern<<<blocks, 256>>>(out, in);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double fmas = fma_per_thread * blocks * 256;
  printf("{\"pattern\": \"%s\", \"ctas_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.2f, \"fma_per_clk_per_sm\": %.1f}\n", name,
         ctas_per_sm, best, 2.0 * fmas / (best * 1e-3) / 1e12,
         fmas / (best * 1e-3) / ((double)clk_khz * 1e3) / p.multiProcessorCount);
}

int main() {
  float* out;
  float4* in;
  cudaMalloc(&out, sizeof(float) * 148 * 8 * 256);
  cudaMalloc(&in, sizeof(float4) * 2048);
  cudaMemset(in, 0, sizeof(float4) * 2048);
  for (int c = 2; c <= 4; c += 2) {
    run("wgrad-like FFMA2", k_wgrad_like<true>, 48.0 * ITERS, c, out, in);
    run("wgrad-like FFMA", k_wgrad_like<false>, 48.0 * ITERS, c, out, in);
    run("fwd-like FFMA2", k_fwd_like<true>, 192.0 * (ITERS / 6 * 6 + 6), c, out, in);
    run("fwd-like FFMA", k_fwd_like<false>, 192.0 * (ITERS / 6 * 6 + 6), c, out, in);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
