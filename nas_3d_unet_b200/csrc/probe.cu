// Roofline probes, timed live by bench.py next to the kernels they bound (the HBM and tensor peaks
// come from MEASURED_PEAKS.json; no measured fp32-FMA figure exists there, so the library carries
// its own probe): packed fp32 FMA chains with no memory traffic - the most the CUDA cores of this
// part deliver to ANY FFMA-bound kernel at the clocks of this run.
#include "common.cuh"

namespace nas3d {

constexpr int kProbeChains = 16;     // independent f32x2 accumulators per thread

__global__ void __launch_bounds__(256) probe_fma_kernel(float* __restrict__ out, int iters, float a, float b) {
  uint64_t acc[kProbeChains];
  uint64_t pa, pb;
  asm("mov.b64 %0, {%1, %1};" : "=l"(pa) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(pb) : "f"(b));
#pragma unroll
  for (int i = 0; i < kProbeChains; ++i) {
    const float lo = (float)(threadIdx.x + 2 * i), hi = (float)(threadIdx.x + 2 * i + 1);
    asm("mov.b64 %0, {%1, %2};" : "=l"(acc[i]) : "f"(lo), "f"(hi));
  }
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < kProbeChains; ++i)
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(pa), "l"(pb));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kProbeChains; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
    s += lo + hi;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace nas3d

using namespace nas3d;

extern "C" {

long long nas3d_probe_fma(float* out, long long out_floats, int iters, void* stream) {
  const int blocks = kNumSMs * 8, threads = 256;
  if (!out || out_floats < (long long)blocks * threads || iters < 1) {
    fail(NAS3D_ERR_ARG, "probe_fma: need %d output floats and iters >= 1", blocks * threads);
    return NAS3D_ERR_ARG;
  }
  probe_fma_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(out, iters, 1.0001f, 0.5f);
  const int rc = launched("probe_fma");
  if (rc) return rc;
  return 2ll * 2 * kProbeChains * (long long)iters * blocks * threads;     // flops of this launch
}

}  // extern "C"
