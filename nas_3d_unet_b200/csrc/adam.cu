// Flat-buffer Adam: one launch updates every parameter of a param group.
// Replaces the per-tensor torch.optim.Adam update of search.py:103-104,228,237 / train.py:58,127.
//
// Parameters, exp_avg and exp_avg_sq live in three flat fp32 arenas (tensor offsets are multiples
// of 4 floats); gradients stay wherever autograd put them and are reached through a table
// of per-tensor pointers passed as kernel parameters (up to kAdamBatch per launch, so nothing is
// uploaded and a captured CUDA graph holds them as node constants).  The work is cut into chunks of at most kChunk floats of ONE tensor;
// each CTA owns one chunk.  HBM-bound: 16 B read + 12 B written per parameter.
//
// Update rule = torch.optim.Adam (amsgrad=False, L2 weight_decay):
//   g' = g + wd*p ; m = m + (1-b1)(g'-m) ; v = b2*v + (1-b2) g'^2
//   p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// lr/betas/eps/wd and the step counter live on the device so a captured CUDA graph sees
// scheduler changes (ReduceLROnPlateau) and keeps counting.
#include "common.cuh"

namespace nas3d {

// hyper[]: 0 lr, 1 beta1, 2 beta2, 3 eps, 4 weight_decay, 5 step (float count), 6 step_size,
//          7 sqrt(bias_correction2), 8 maximize, 9 (1-beta1), 10 (1-beta2) rounded from double by the
//          host exactly as torch does
__global__ void adam_tick_kernel(float* hyper) {
  const double step = (double)hyper[5] + 1.0;
  hyper[5] = (float)step;
  const double bc1 = 1.0 - pow((double)hyper[1], step);
  const double bc2 = 1.0 - pow((double)hyper[2], step);
  hyper[6] = (float)((double)hyper[0] / bc1);
  hyper[7] = (float)sqrt(bc2);
}

constexpr int kAdamTB = 256;
constexpr int kAdamBatch = 1024;   // gradient pointers per launch (8 KB of the 32 KB parameter space)

struct AdamPtrs {
  const float* g[kAdamBatch];
};

__device__ __forceinline__ void adam_one(float& p, float& m, float& v, float g, float omb1,
                                         float b2, float omb2, float eps, float wd,
                                         float step_size, float bc2s) {
  g = fmaf(wd, p, g);
  m = fmaf(omb1, g - m, m);
  v = fmaf(omb2, g * g, b2 * v);
  const float denom = sqrtf(v) / bc2s + eps;
  p = p - step_size * (m / denom);
}

__global__ void __launch_bounds__(kAdamTB)
    adam_flat_kernel(float* __restrict__ param, float* __restrict__ exp_avg,
                     float* __restrict__ exp_avg_sq, const __grid_constant__ AdamPtrs ptrs,
                     int tensor0, const int4* __restrict__ chunks,
                     const float* __restrict__ hyper) {
  const int4 ch = chunks[blockIdx.x];   // tensor, flat offset, offset inside the tensor, length
  const float* g = ptrs.g[ch.x - tensor0];
  if (g == nullptr) return;             // parameter without a gradient this step: untouched
  g += ch.z;
  const float b2 = hyper[2], eps = hyper[3], wd = hyper[4];
  const float omb1 = hyper[9], omb2 = hyper[10];
  const float step_size = hyper[6], bc2s = hyper[7];
  const float sign = hyper[8] != 0.f ? -1.f : 1.f;
  float* p = param + ch.y;
  float* m = exp_avg + ch.y;
  float* v = exp_avg_sq + ch.y;
  const int len = ch.w;
  const int n4 = ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) ? (len >> 2) : 0;
  for (int i = threadIdx.x; i < n4; i += kAdamTB) {
    float4 pp = *reinterpret_cast<float4*>(p + 4 * i);
    float4 mm = *reinterpret_cast<float4*>(m + 4 * i);
    float4 vv = *reinterpret_cast<float4*>(v + 4 * i);
    const float4 gg = ldg4(g + 4 * i);
    adam_one(pp.x, mm.x, vv.x, sign * gg.x, omb1, b2, omb2, eps, wd, step_size, bc2s);
    adam_one(pp.y, mm.y, vv.y, sign * gg.y, omb1, b2, omb2, eps, wd, step_size, bc2s);
    adam_one(pp.z, mm.z, vv.z, sign * gg.z, omb1, b2, omb2, eps, wd, step_size, bc2s);
    adam_one(pp.w, mm.w, vv.w, sign * gg.w, omb1, b2, omb2, eps, wd, step_size, bc2s);
    st4(p + 4 * i, pp);
    st4(m + 4 * i, mm);
    st4(v + 4 * i, vv);
  }
  for (int i = 4 * n4 + threadIdx.x; i < len; i += kAdamTB) {
    float pp = p[i], mm = m[i], vv = v[i];
    adam_one(pp, mm, vv, sign * g[i], omb1, b2, omb2, eps, wd, step_size, bc2s);
    p[i] = pp;
    m[i] = mm;
    v[i] = vv;
  }
}

}  // namespace nas3d

using namespace nas3d;

extern "C" {

int nas3d_adam_chunk_floats(void) { return 4096; }

int nas3d_adam_flat_step(float* param, float* exp_avg, float* exp_avg_sq,
                         const float* const* grad_ptrs, int ntensors, const int* tensor_first_chunk,
                         const int* chunks_dev, float* hyper_dev, void* stream) {
  NAS3D_REQUIRE(param && exp_avg && exp_avg_sq && hyper_dev, "adam_flat_step: null pointer");
  NAS3D_REQUIRE(ntensors >= 0 && (ntensors == 0 || (grad_ptrs && tensor_first_chunk && chunks_dev)),
                "adam_flat_step: bad tensor table");
  NAS3D_REQUIRE(aligned16(param) && aligned16(exp_avg) && aligned16(exp_avg_sq) &&
                    aligned16(chunks_dev),
                "adam_flat_step: arenas and the chunk table must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  adam_tick_kernel<<<1, 1, 0, st>>>(hyper_dev);
  int rc = launched("adam_tick");
  if (rc != NAS3D_OK) return rc;
  const int4* chunks = reinterpret_cast<const int4*>(chunks_dev);
  for (int t0 = 0; t0 < ntensors; t0 += kAdamBatch) {
    const int t1 = t0 + kAdamBatch < ntensors ? t0 + kAdamBatch : ntensors;
    const int c0 = tensor_first_chunk[t0], c1 = tensor_first_chunk[t1];
    NAS3D_REQUIRE(c1 >= c0, "adam_flat_step: tensor_first_chunk must be non-decreasing");
    if (c1 == c0) continue;
    AdamPtrs P;
    bool any = false;
    for (int t = t0; t < t1; ++t) {
      P.g[t - t0] = grad_ptrs[t];
      any = any || grad_ptrs[t] != nullptr;
    }
    if (!any) continue;
    adam_flat_kernel<<<c1 - c0, kAdamTB, 0, st>>>(param, exp_avg, exp_avg_sq, P, t0, chunks + c0,
                                                  hyper_dev);
    rc = launched("adam_flat");
    if (rc != NAS3D_OK) return rc;
  }
  return NAS3D_OK;
}

}  // extern "C"
