// Fused affine-sum kernels: the MixedOp / node-sum / GroupNorm-apply / ReLU / SE-scale /
// concat hot loop of cell.py:30,32,81,82 and prim_ops.py:75-80,152 as ONE pass over memory.
//
//   fwd : out = sum_k w_k * act_k(a_k[n,c]*x_k + b_k[n,c])          reads K tensors, writes 1
//   bwd : dx_k (+)= p_k*m_k*dout + q_k*x_k + r_k                    reads dout (+x_k), writes K
//
// HBM-bound: one float4 (4 channels of one voxel) per thread per iteration, grid-stride,
// grid = 148 SMs x 8 CTAs.  Per-(n,c) coefficients are tiny and stay in L1.
#include <stdlib.h>
#include "common.cuh"
#include "gn_coef.cuh"

namespace nas3d {

struct FwdTerms {
  const float* x[NAS3D_MAX_TERMS];
  const float* a[NAS3D_MAX_TERMS];
  const float* b[NAS3D_MAX_TERMS];
  const float* w[NAS3D_MAX_TERMS];
  int ld[NAS3D_MAX_TERMS];
  int relu[NAS3D_MAX_TERMS];
  int nterms;
};

// coefficient load: the read-only path, or a plain (coherent) load when the coefficients were
// written by this very kernel's prologue
template <bool COH>
__device__ __forceinline__ float4 coef4(const float* p) {
  return COH ? *reinterpret_cast<const float4*>(p) : ldg4(p);
}

// grid = (chunks, N): blockIdx.y is the sample, so no 64-bit division in the hot loop; every
// thread keeps two independent float4 elements (x K terms) in flight.
template <bool COH>
__device__ __forceinline__ float4 affine_term(const FwdTerms& T, int k, const float* base_k,
                                              unsigned vox, int c, long long nc) {
  float4 v = ldg4(base_k + (long long)vox * T.ld[k] + c);
  if (T.a[k]) {
    const float4 a = coef4<COH>(T.a[k] + nc);
    v.x *= a.x; v.y *= a.y; v.z *= a.z; v.w *= a.w;
  }
  if (T.b[k]) {
    const float4 b = coef4<COH>(T.b[k] + nc);
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
  }
  if (T.relu[k]) {
    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
  }
  return v;
}

template <bool COH>
__device__ __forceinline__ void affine_sum_fwd_loop(const FwdTerms& T, float* __restrict__ out,
                                                    int ld_out, long long V, int C, int C4,
                                                    unsigned per_sample) {
  const int n = blockIdx.y;
  const long long vbase = (long long)n * V;
  float* outn = out + vbase * ld_out;
  const unsigned stride = gridDim.x * 256u;
  for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < per_sample; i += 2 * stride) {
    const unsigned i2 = i + stride;
    const bool has2 = i2 < per_sample;
    const unsigned vox0 = i / (unsigned)C4, vox1 = has2 ? i2 / (unsigned)C4 : vox0;
    const int c0 = (int)(i - vox0 * C4) * 4, c1 = has2 ? (int)(i2 - vox1 * C4) * 4 : c0;
    const long long nc0 = (long long)n * C + c0, nc1 = (long long)n * C + c1;
    float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
#pragma unroll 2
    for (int k = 0; k < T.nterms; ++k) {
      const float* bk = T.x[k] + vbase * T.ld[k];
      const float4 v0 = affine_term<COH>(T, k, bk, vox0, c0, nc0);
      const float4 v1 = affine_term<COH>(T, k, bk, vox1, c1, nc1);
      const float w = T.w[k] ? __ldg(T.w[k]) : 1.f;
      acc0.x += w * v0.x; acc0.y += w * v0.y; acc0.z += w * v0.z; acc0.w += w * v0.w;
      acc1.x += w * v1.x; acc1.y += w * v1.y; acc1.z += w * v1.z; acc1.w += w * v1.w;
    }
    st4(outn + (long long)vox0 * ld_out + c0, acc0);
    if (has2) st4(outn + (long long)vox1 * ld_out + c1, acc1);
  }
}

__global__ void __launch_bounds__(256)
    affine_sum_fwd_kernel(const __grid_constant__ FwdTerms T, float* __restrict__ out, int ld_out,
                          long long V, int C, int C4, unsigned per_sample) {
  affine_sum_fwd_loop<false>(T, out, ld_out, V, C, C4, per_sample);
}

// Same pass with the GroupNorm coefficient kernels of its terms folded in: every CTA first derives
// a, b (and mean / rstd, kept for the backward) of its sample from the fp64 moments - a few hundred
// flops - instead of waiting for a separate 5 us launch per node on the critical path.  All CTAs of
// a sample store identical values; the loop then reads them with coherent loads.
__global__ void __launch_bounds__(256)
    affine_sum_fwd_gn_kernel(const __grid_constant__ FwdTerms T, const __grid_constant__ GnFwdBatch B,
                             int G, double inv_m, float eps, float* __restrict__ out, int ld_out,
                             long long V, int C, int C4, unsigned per_sample) {
  __shared__ GnScratch sc;
  for (int k = 0; k < T.nterms; ++k)
    if (B.S[k]) gn_coef_body(B, k, blockIdx.y, C, G, inv_m, eps, sc);
  __syncthreads();
  affine_sum_fwd_loop<true>(T, out, ld_out, V, C, C4, per_sample);
}

// Ring-staged forward for the big few-term sums (K <= 3): same element mapping and arithmetic, but
// a thread's next AR_S-1 elements of every term are in flight as cp.async copies into its own
// slots of a shared-memory ring (see reduce.cu / conv_pointwise_bwd.cu: the register-staged
// streaming kernels are latency-bound at 5.0-5.2 TB/s).  gridDim.x * 256 is a multiple of C4, so
// a thread's channel group is fixed and every address is a running pointer.
constexpr int AR_S = 5;

__device__ __forceinline__ void ar_cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(
                   (unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

template <int K>
__global__ void __launch_bounds__(256)
    affine_sum_fwd_ring_kernel(const __grid_constant__ FwdTerms T, float* __restrict__ out,
                               int ld_out, long long V, int C, int C4, unsigned per_sample) {
  extern __shared__ __align__(16) float4 aring[];     // [term][stage][thread]
  const int n = blockIdx.y;
  const unsigned stride = gridDim.x * 256u;
  const unsigned i0 = blockIdx.x * 256u + threadIdx.x;
  const unsigned vox0 = i0 / (unsigned)C4;
  const int c = (int)(i0 - vox0 * C4) * 4;
  const unsigned vs = stride / (unsigned)C4;
  const int trips = i0 < per_sample ? (int)((per_sample - i0 + stride - 1) / stride) : 0;
  const long long nc = (long long)n * C + c;

  const float* px[K];
  long long sx[K];
  float4 a[K], b[K];
  float w[K];
  bool relu[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    px[k] = T.x[k] + ((long long)n * V + vox0) * T.ld[k] + c;
    sx[k] = (long long)vs * T.ld[k];
    a[k] = T.a[k] ? ldg4(T.a[k] + nc) : make_float4(1.f, 1.f, 1.f, 1.f);
    b[k] = T.b[k] ? ldg4(T.b[k] + nc) : make_float4(0.f, 0.f, 0.f, 0.f);
    w[k] = T.w[k] ? __ldg(T.w[k]) : 1.f;
    relu[k] = T.relu[k] != 0;
  }
  float* po = out + ((long long)n * V + vox0) * ld_out + c;
  const long long so = (long long)vs * ld_out;

  float4* const r0 = aring + threadIdx.x;
  float4* const r_end = r0 + AR_S * 256;
  float4* is = r0;
  int issued = 0;
  auto issue = [&]() {
    if (issued < trips) {
#pragma unroll
      for (int k = 0; k < K; ++k) ar_cp16(is + k * AR_S * 256, px[k]);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    ++issued;
#pragma unroll
    for (int k = 0; k < K; ++k) px[k] += sx[k];
    is += 256;
    if (is == r_end) is = r0;
  };
#pragma unroll
  for (int s = 0; s < AR_S - 1; ++s) issue();
  const float4* cs = r0;
#pragma unroll 1
  for (int it = 0; it < trips; ++it) {
    issue();
    asm volatile("cp.async.wait_group %0;\n" ::"n"(AR_S - 1) : "memory");
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float4 v = cs[k * AR_S * 256];
      v.x *= a[k].x; v.y *= a[k].y; v.z *= a[k].z; v.w *= a[k].w;
      v.x += b[k].x; v.y += b[k].y; v.z += b[k].z; v.w += b[k].w;
      if (relu[k]) {
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
      }
      acc.x += w[k] * v.x; acc.y += w[k] * v.y; acc.z += w[k] * v.z; acc.w += w[k] * v.w;
    }
    st4(po, acc);
    po += so;
    cs += 256;
    if (cs == r_end) cs = r0;
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

template <int K>
static int launch_fwd_ring(const FwdTerms& T, float* out, int ld_out, int N, long long V, int C,
                           int C4, long long per_sample, cudaStream_t st) {
  auto kern = affine_sum_fwd_ring_kernel<K>;
  const int smem = K * AR_S * 256 * (int)sizeof(float4);
  static int resident = -1;
  if (resident < 0) {
    NAS3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int o = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, 256, smem) != cudaSuccess || o < 1) o = 2;
    resident = o * kNumSMs;
  }
  // one resident wave; gridDim.x * 256 must be a multiple of C4 (C4 is 3 or a power of two <= 256)
  long long bx = resident / N;
  const long long need = (per_sample + 255) / 256;
  if (bx > need) bx = need;
  if (C4 % 3 == 0) bx -= bx % 3;
  if (bx < 1 || (bx * 256) % C4 != 0) return NAS3D_ERR_UNSUPPORTED;
  kern<<<dim3((unsigned)bx, (unsigned)N), 256, smem, st>>>(T, out, ld_out, V, C, C4, (unsigned)per_sample);
  return NAS3D_OK;
}

struct BwdTerms {
  const float* x[NAS3D_MAX_TERMS];
  const float* a[NAS3D_MAX_TERMS];
  const float* b[NAS3D_MAX_TERMS];
  const float* p[NAS3D_MAX_TERMS];
  const float* q[NAS3D_MAX_TERMS];
  const float* r[NAS3D_MAX_TERMS];
  const float* w[NAS3D_MAX_TERMS];
  float* dx[NAS3D_MAX_TERMS];
  int ld[NAS3D_MAX_TERMS];
  int ld_dx[NAS3D_MAX_TERMS];
  int relu[NAS3D_MAX_TERMS];
  int acc[NAS3D_MAX_TERMS];
  int nterms;
};

// dx_k (+)= p_k*m_k*dout + q_k*x_k + r_k for one float4 element of term k
template <bool COH>
__device__ __forceinline__ void bwd_apply_one(const BwdTerms& T, int k, float4 g, const float4& x,
                                              long long nc, float* dst) {
  if (T.relu[k]) {
    float4 a = make_float4(1.f, 1.f, 1.f, 1.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (T.a[k]) a = ldg4(T.a[k] + nc);
    if (T.b[k]) b = ldg4(T.b[k] + nc);
    g.x = (a.x * x.x + b.x > 0.f) ? g.x : 0.f;
    g.y = (a.y * x.y + b.y > 0.f) ? g.y : 0.f;
    g.z = (a.z * x.z + b.z > 0.f) ? g.z : 0.f;
    g.w = (a.w * x.w + b.w > 0.f) ? g.w : 0.f;
  }
  if (T.p[k]) {
    const float4 p = coef4<COH>(T.p[k] + nc);
    g.x *= p.x; g.y *= p.y; g.z *= p.z; g.w *= p.w;
  } else if (T.w[k]) {
    const float w = __ldg(T.w[k]);
    g.x *= w; g.y *= w; g.z *= w; g.w *= w;
  }
  if (T.q[k]) {
    const float4 q = coef4<COH>(T.q[k] + nc);
    g.x += q.x * x.x; g.y += q.y * x.y; g.z += q.z * x.z; g.w += q.w * x.w;
  }
  if (T.r[k]) {
    const float4 r = coef4<COH>(T.r[k] + nc);
    g.x += r.x; g.y += r.y; g.z += r.z; g.w += r.w;
  }
  if (T.acc[k]) {
    // plain load (not the read-only path): an earlier term of this very launch may have
    // written this location from this same thread
    const float4 o = *reinterpret_cast<const float4*>(dst);
    g.x += o.x; g.y += o.y; g.z += o.z; g.w += o.w;
  }
  st4(dst, g);
}

// Two independent float4 elements per thread per iteration (as in the forward kernel): the dout
// and x_k loads of both are issued before the first dependent use, doubling the bytes in flight
// per thread (r1d ncu: 4.6 TB/s with one element in flight per thread, latency-bound).
template <bool COH>
__device__ __forceinline__ void affine_sum_bwd_apply_loop(const BwdTerms& T,
                                                          const float* __restrict__ dout,
                                                          int ld_dout, long long V, int C, int C4,
                                                          unsigned per_sample) {
  const int n = blockIdx.y;
  const long long vbase = (long long)n * V;
  const float* dn = dout + vbase * ld_dout;
  const unsigned stride = gridDim.x * 256u;
  for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < per_sample; i += 2 * stride) {
    const unsigned i2 = i + stride;
    const bool has2 = i2 < per_sample;
    const unsigned vox0 = i / (unsigned)C4, vox1 = has2 ? i2 / (unsigned)C4 : vox0;
    const int c0 = (int)(i - vox0 * C4) * 4, c1 = has2 ? (int)(i2 - vox1 * C4) * 4 : c0;
    const long long nc0 = (long long)n * C + c0, nc1 = (long long)n * C + c1;
    const float4 d0 = ldg4(dn + (long long)vox0 * ld_dout + c0);
    const float4 d1 = ldg4(dn + (long long)vox1 * ld_dout + c1);
    for (int k = 0; k < T.nterms; ++k) {
      const bool need_x = T.relu[k] || T.q[k];
      float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
      if (need_x) {
        x0 = ldg4(T.x[k] + (vbase + vox0) * T.ld[k] + c0);
        x1 = ldg4(T.x[k] + (vbase + vox1) * T.ld[k] + c1);
      }
      bwd_apply_one<COH>(T, k, d0, x0, nc0, T.dx[k] + (vbase + vox0) * T.ld_dx[k] + c0);
      if (has2) bwd_apply_one<COH>(T, k, d1, x1, nc1, T.dx[k] + (vbase + vox1) * T.ld_dx[k] + c1);
    }
  }
}

__global__ void __launch_bounds__(256)
    affine_sum_bwd_apply_kernel(const __grid_constant__ BwdTerms T, const float* __restrict__ dout,
                                int ld_dout, long long V, int C, int C4, unsigned per_sample) {
  affine_sum_bwd_apply_loop<false>(T, dout, ld_dout, V, C, C4, per_sample);
}

// With the GroupNorm backward coefficient kernel of its terms folded in (see the forward): every
// CTA derives p, q, r of its sample from the reductions R; the CTA with blockIdx.x == 0 also adds
// the parameter gradients that fall out of the same sums (d gamma, d beta, the bias of the conv
// that fed the GroupNorm, d alpha).
__global__ void __launch_bounds__(256)
    affine_sum_bwd_apply_gn_kernel(const __grid_constant__ BwdTerms T,
                                   const __grid_constant__ GnBwdBatch B, int G, double inv_m,
                                   const float* __restrict__ dout, int ld_dout, long long V, int C,
                                   int C4, unsigned per_sample) {
  __shared__ GnScratch sc;
  for (int k = 0; k < T.nterms; ++k)
    if (B.R[k]) gn_bwd_coef_body(B, k, blockIdx.y, C, G, inv_m, (double)V, blockIdx.x == 0, sc);
  __syncthreads();
  affine_sum_bwd_apply_loop<true>(T, dout, ld_dout, V, C, C4, per_sample);
}

// Ring-staged backward apply for big sums of <= 2 terms (same staging as affine_sum_fwd_ring_kernel;
// NOT YET MEASURED ON A B200 - off unless NAS3D_APPLY_RING=1).  Ring
// components: dout | x_k of the terms that need x (relu mask or q) | old dx_k of the accumulating
// terms.  The host only takes this path when no two terms share a dx buffer.
template <int K>
__global__ void __launch_bounds__(256)
    affine_sum_bwd_apply_ring_kernel(const __grid_constant__ BwdTerms T, const float* __restrict__ dout,
                                     int ld_dout, long long V, int C, int C4, unsigned per_sample) {
  extern __shared__ __align__(16) float4 aring[];     // [component][stage][thread]
  const int n = blockIdx.y;
  const unsigned stride = gridDim.x * 256u;           // multiple of C4 (host)
  const unsigned i0 = blockIdx.x * 256u + threadIdx.x;
  const unsigned vox0 = i0 / (unsigned)C4;
  const int c = (int)(i0 - vox0 * C4) * 4;
  const unsigned vs = stride / (unsigned)C4;
  const int trips = i0 < per_sample ? (int)((per_sample - i0 + stride - 1) / stride) : 0;
  const long long nc = (long long)n * C + c;

  const float* pd = dout + ((long long)n * V + vox0) * ld_dout + c;
  const long long sd = (long long)vs * ld_dout;
  const float* px[K];
  float* pdx[K];
  long long sx[K], sdx[K];
  bool need_x[K], relu[K], acc[K], has_q[K], has_r[K];
  int cx[K], cold[K];                                 // ring component of x_k / old dx_k
  float4 a[K], b[K], pc[K], q[K], r[K];
  int ncomp = 1;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    relu[k] = T.relu[k] != 0;
    has_q[k] = T.q[k] != nullptr;
    has_r[k] = T.r[k] != nullptr;
    need_x[k] = relu[k] || has_q[k];
    acc[k] = T.acc[k] != 0;
    px[k] = need_x[k] ? T.x[k] + ((long long)n * V + vox0) * T.ld[k] + c : nullptr;
    sx[k] = (long long)vs * T.ld[k];
    pdx[k] = T.dx[k] + ((long long)n * V + vox0) * T.ld_dx[k] + c;
    sdx[k] = (long long)vs * T.ld_dx[k];
    a[k] = (relu[k] && T.a[k]) ? ldg4(T.a[k] + nc) : make_float4(1.f, 1.f, 1.f, 1.f);
    b[k] = (relu[k] && T.b[k]) ? ldg4(T.b[k] + nc) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (T.p[k]) {
      pc[k] = ldg4(T.p[k] + nc);
    } else {
      const float w = T.w[k] ? __ldg(T.w[k]) : 1.f;
      pc[k] = make_float4(w, w, w, w);
    }
    q[k] = has_q[k] ? ldg4(T.q[k] + nc) : make_float4(0.f, 0.f, 0.f, 0.f);
    r[k] = has_r[k] ? ldg4(T.r[k] + nc) : make_float4(0.f, 0.f, 0.f, 0.f);
    cx[k] = need_x[k] ? ncomp++ : 0;
  }
#pragma unroll
  for (int k = 0; k < K; ++k) cold[k] = acc[k] ? ncomp++ : 0;

  float4* const r0 = aring + threadIdx.x;
  float4* const r_end = r0 + AR_S * 256;
  float4* is = r0;
  int issued = 0;
  const float* pold[K];
#pragma unroll
  for (int k = 0; k < K; ++k) pold[k] = pdx[k];
  auto issue = [&]() {
    if (issued < trips) {
      ar_cp16(is, pd);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (need_x[k]) ar_cp16(is + cx[k] * AR_S * 256, px[k]);
        if (acc[k]) ar_cp16(is + cold[k] * AR_S * 256, pold[k]);
      }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    ++issued;
    pd += sd;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (need_x[k]) px[k] += sx[k];
      pold[k] += sdx[k];
    }
    is += 256;
    if (is == r_end) is = r0;
  };
#pragma unroll
  for (int s = 0; s < AR_S - 1; ++s) issue();
  const float4* cs = r0;
#pragma unroll 1
  for (int it = 0; it < trips; ++it) {
    issue();
    asm volatile("cp.async.wait_group %0;\n" ::"n"(AR_S - 1) : "memory");
    const float4 d = cs[0];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float4 g = d;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (need_x[k]) x = cs[cx[k] * AR_S * 256];
      if (relu[k]) {
        g.x = (a[k].x * x.x + b[k].x > 0.f) ? g.x : 0.f;
        g.y = (a[k].y * x.y + b[k].y > 0.f) ? g.y : 0.f;
        g.z = (a[k].z * x.z + b[k].z > 0.f) ? g.z : 0.f;
        g.w = (a[k].w * x.w + b[k].w > 0.f) ? g.w : 0.f;
      }
      g.x *= pc[k].x; g.y *= pc[k].y; g.z *= pc[k].z; g.w *= pc[k].w;
      if (has_q[k]) { g.x += q[k].x * x.x; g.y += q[k].y * x.y; g.z += q[k].z * x.z; g.w += q[k].w * x.w; }
      if (has_r[k]) { g.x += r[k].x; g.y += r[k].y; g.z += r[k].z; g.w += r[k].w; }
      if (acc[k]) {
        const float4 o = cs[cold[k] * AR_S * 256];
        g.x += o.x; g.y += o.y; g.z += o.z; g.w += o.w;
      }
      st4(pdx[k], g);
      pdx[k] += sdx[k];
    }
    cs += 256;
    if (cs == r_end) cs = r0;
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

template <int K>
static int launch_bwd_apply_ring(const BwdTerms& T, const float* dout, int ld_dout, int N, long long V,
                                 int C, int C4, long long per_sample, cudaStream_t st) {
  auto kern = affine_sum_bwd_apply_ring_kernel<K>;
  int ncomp = 1;
  for (int k = 0; k < K; ++k) ncomp += ((T.relu[k] || T.q[k]) ? 1 : 0) + (T.acc[k] ? 1 : 0);
  const int smem = ncomp * AR_S * 256 * (int)sizeof(float4);
  static int occ[2 * K + 2] = {0};     // by component count
  if (occ[ncomp] == 0) {
    NAS3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (2 * K + 1) * AR_S * 256 * (int)sizeof(float4)));
    int o = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, 256, smem) != cudaSuccess || o < 1) o = 1;
    occ[ncomp] = o;
  }
  long long bx = (long long)occ[ncomp] * kNumSMs / N;
  const long long need = (per_sample + 255) / 256;
  if (bx > need) bx = need;
  if (C4 % 3 == 0) bx -= bx % 3;
  if (bx < 1 || (bx * 256) % C4 != 0) return NAS3D_ERR_UNSUPPORTED;
  kern<<<dim3((unsigned)bx, (unsigned)N), 256, smem, st>>>(T, dout, ld_dout, V, C, C4, (unsigned)per_sample);
  return NAS3D_OK;
}

__global__ void __launch_bounds__(256)
    add_inplace_kernel(float* __restrict__ y, const float* __restrict__ x, long long n4,
                       long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 a = *reinterpret_cast<float4*>(y + i * 4);
    const float4 b = ldg4(x + i * 4);
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    st4(y + i * 4, a);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) y[n4 * 4 + threadIdx.x] += x[n4 * 4 + threadIdx.x];
}

__global__ void __launch_bounds__(256)
    sigmoid_bwd_kernel(const float* __restrict__ prob, const float* __restrict__ dprob,
                       float* __restrict__ dlogit, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float p = __ldg(prob + i);
    dlogit[i] = __ldg(dprob + i) * p * (1.f - p);
  }
}

// NCDHW -> NDHWC: one thread per voxel, coalesced plane reads, C-wide row writes
__global__ void __launch_bounds__(256)
    ncdhw_to_ndhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int C,
                          long long V, int ld, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / V, v = i - n * V;
    const float* s = src + n * C * V + v;
    float* d = dst + i * ld;
    if ((C & 3) == 0 && (ld & 3) == 0) {
      for (int c = 0; c < C; c += 4) {
        float4 t;
        t.x = __ldg(s + (long long)(c + 0) * V);
        t.y = __ldg(s + (long long)(c + 1) * V);
        t.z = __ldg(s + (long long)(c + 2) * V);
        t.w = __ldg(s + (long long)(c + 3) * V);
        st4(d + c, t);
      }
    } else {
      for (int c = 0; c < C; ++c) d[c] = __ldg(s + (long long)c * V);
    }
  }
}

static inline unsigned grid_for(long long total, int block) {
  long long b = (total + block - 1) / block;
  long long cap = (long long)kNumSMs * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

// 2-D grid: x covers one sample (elems_per_block elements per CTA pass), y = sample; the x extent
// is capped so that the whole grid is ~8 CTAs per SM
static inline dim3 grid2d(long long per_sample, int N, int elems_per_block) {
  long long bx = (per_sample + elems_per_block - 1) / elems_per_block;
  long long cap = ((long long)kNumSMs * 8 + N - 1) / N;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  return dim3((unsigned)bx, (unsigned)N);
}

}  // namespace nas3d

using namespace nas3d;

extern "C" {

int nas3d_affine_sum_fwd(int nterms, const float* const* x, const int* ld_x,
                         const float* const* a, const float* const* b, const float* const* w,
                         const int* relu, float* out, int ld_out, int N, long long V, int C,
                         void* stream) {
  NAS3D_REQUIRE(nterms >= 1 && nterms <= NAS3D_MAX_TERMS, "affine_sum_fwd: nterms=%d", nterms);
  NAS3D_REQUIRE(C % 4 == 0 && ld_out % 4 == 0 && aligned16(out),
                "affine_sum_fwd: C=%d ld_out=%d must be multiples of 4, out 16B aligned", C, ld_out);
  FwdTerms T;
  T.nterms = nterms;
  for (int k = 0; k < nterms; ++k) {
    NAS3D_REQUIRE(ld_x[k] % 4 == 0 && aligned16(x[k]), "affine_sum_fwd: term %d misaligned", k);
    T.x[k] = x[k]; T.ld[k] = ld_x[k];
    T.a[k] = a ? a[k] : nullptr; T.b[k] = b ? b[k] : nullptr; T.w[k] = w ? w[k] : nullptr;
    T.relu[k] = relu ? relu[k] : 0;
  }
  const int C4 = C / 4;
  const long long per_sample = V * C4;
  NAS3D_REQUIRE(per_sample < (1ll << 31), "affine_sum_fwd: sample too large for 32-bit indexing");
  // ring-staged kernel for big sums of <= 3 terms (option affine_ring = 0: the register-staged
  // kernel).  Measured on B200: affine_sum_fwd 1.62 -> 1.50 ms per searched-net step, 420.7 -> 426.4
  // patches/s (profiles/r1f_ab_affine_ring.json)
  if (g_opt.affine_ring && nterms <= 3 && per_sample * N >= ring_min_elems() &&
      (C4 == 3 || (C4 & (C4 - 1)) == 0) && C4 <= 256) {
    int rc = nterms == 1 ? launch_fwd_ring<1>(T, out, ld_out, N, V, C, C4, per_sample, (cudaStream_t)stream)
           : nterms == 2 ? launch_fwd_ring<2>(T, out, ld_out, N, V, C, C4, per_sample, (cudaStream_t)stream)
                         : launch_fwd_ring<3>(T, out, ld_out, N, V, C, C4, per_sample, (cudaStream_t)stream);
    if (rc == NAS3D_OK) return launched("affine_sum_fwd_ring");
    if (rc != NAS3D_ERR_UNSUPPORTED) return rc;
  }
  affine_sum_fwd_kernel<<<grid2d(per_sample, N, 512), 256, 0, (cudaStream_t)stream>>>(
      T, out, ld_out, V, C, C4, (unsigned)per_sample);
  return launched("affine_sum_fwd");
}

int nas3d_affine_sum_fwd_gn(int nterms, const float* const* x, const int* ld_x,
                            float* const* a, float* const* b, const float* const* w,
                            const int* relu, const double* const* gn_S,
                            const float* const* gn_gamma, const float* const* gn_beta,
                            float* const* gn_mean_rstd, int G, float eps, float* out, int ld_out,
                            int N, long long V, int C, void* stream) {
  NAS3D_REQUIRE(nterms >= 1 && nterms <= NAS3D_MAX_TERMS, "affine_sum_fwd_gn: nterms=%d", nterms);
  NAS3D_REQUIRE(C % 4 == 0 && ld_out % 4 == 0 && aligned16(out),
                "affine_sum_fwd_gn: C=%d ld_out=%d must be multiples of 4, out 16B aligned", C, ld_out);
  NAS3D_REQUIRE(G >= 1 && G <= 64 && C % G == 0, "affine_sum_fwd_gn: bad groups %d for C=%d", G, C);
  NAS3D_REQUIRE(gn_S && gn_gamma && gn_beta && gn_mean_rstd && a && b, "affine_sum_fwd_gn: NULL array");
  FwdTerms T;
  GnFwdBatch B;
  T.nterms = nterms;
  for (int k = 0; k < nterms; ++k) {
    NAS3D_REQUIRE(ld_x[k] % 4 == 0 && aligned16(x[k]), "affine_sum_fwd_gn: term %d misaligned", k);
    T.x[k] = x[k]; T.ld[k] = ld_x[k];
    T.a[k] = a[k]; T.b[k] = b[k]; T.w[k] = w ? w[k] : nullptr;
    T.relu[k] = relu ? relu[k] : 0;
    B.S[k] = gn_S[k]; B.gamma[k] = gn_gamma[k]; B.beta[k] = gn_beta[k];
    B.a[k] = a[k]; B.b[k] = b[k]; B.mr[k] = gn_mean_rstd[k];
    NAS3D_REQUIRE(!B.S[k] || (B.gamma[k] && B.beta[k] && B.a[k] && B.b[k] && B.mr[k]),
                  "affine_sum_fwd_gn: GroupNorm term %d needs gamma, beta, a, b, mean_rstd", k);
  }
  const int C4 = C / 4;
  const long long per_sample = V * C4;
  NAS3D_REQUIRE(per_sample < (1ll << 31), "affine_sum_fwd_gn: sample too large for 32-bit indexing");
  const double inv_m = 1.0 / ((double)(C / G) * (double)V);
  affine_sum_fwd_gn_kernel<<<grid2d(per_sample, N, 512), 256, 0, (cudaStream_t)stream>>>(
      T, B, G, inv_m, eps, out, ld_out, V, C, C4, (unsigned)per_sample);
  return launched("affine_sum_fwd_gn");
}

int nas3d_affine_sum_bwd_apply_gn(int nterms, const float* const* x, const int* ld_x,
                                  const float* const* a, const float* const* b, const int* relu,
                                  float* const* p, float* const* q, float* const* r,
                                  const float* const* w, float* const* dx, const int* ld_dx,
                                  const int* accumulate, const float* dout, int ld_dout,
                                  const double* const* gn_R, const float* const* gn_mean_rstd,
                                  const float* const* gn_gamma, float* const* gn_dgamma,
                                  float* const* gn_dbeta, float* const* gn_dw,
                                  const double* const* gn_S, float* const* gn_dbias_prev, int G,
                                  int N, long long V, int C, void* stream) {
  NAS3D_REQUIRE(nterms >= 1 && nterms <= NAS3D_MAX_TERMS, "affine_sum_bwd_apply_gn: nterms=%d", nterms);
  NAS3D_REQUIRE(C % 4 == 0 && ld_dout % 4 == 0 && aligned16(dout),
                "affine_sum_bwd_apply_gn: C=%d ld_dout=%d", C, ld_dout);
  NAS3D_REQUIRE(G >= 1 && G <= 64 && C % G == 0, "affine_sum_bwd_apply_gn: bad groups %d for C=%d", G, C);
  NAS3D_REQUIRE(gn_R && gn_mean_rstd && gn_gamma && gn_dgamma && gn_dbeta && p && q && r,
                "affine_sum_bwd_apply_gn: NULL array");
  BwdTerms T;
  GnBwdBatch B;
  T.nterms = nterms;
  for (int k = 0; k < nterms; ++k) {
    NAS3D_REQUIRE(ld_dx[k] % 4 == 0 && aligned16(dx[k]), "affine_sum_bwd_apply_gn: dx %d misaligned", k);
    T.x[k] = x ? x[k] : nullptr; T.ld[k] = ld_x ? ld_x[k] : 0;
    T.a[k] = a ? a[k] : nullptr; T.b[k] = b ? b[k] : nullptr;
    T.p[k] = p[k]; T.q[k] = q[k]; T.r[k] = r[k];
    T.w[k] = w ? w[k] : nullptr;
    T.dx[k] = dx[k]; T.ld_dx[k] = ld_dx[k];
    T.relu[k] = relu ? relu[k] : 0; T.acc[k] = accumulate ? accumulate[k] : 0;
    NAS3D_REQUIRE(!(T.relu[k] || T.q[k]) || T.x[k], "affine_sum_bwd_apply_gn: term %d needs x", k);
    B.R[k] = gn_R[k]; B.mr[k] = gn_mean_rstd[k]; B.gamma[k] = gn_gamma[k];
    B.a[k] = T.a[k]; B.b[k] = T.b[k]; B.w[k] = T.w[k];
    B.p[k] = p[k]; B.q[k] = q[k]; B.r[k] = r[k];
    B.dgamma[k] = gn_dgamma[k]; B.dbeta[k] = gn_dbeta[k]; B.dw[k] = gn_dw ? gn_dw[k] : nullptr;
    B.S[k] = gn_S ? gn_S[k] : nullptr; B.dbias[k] = gn_dbias_prev ? gn_dbias_prev[k] : nullptr;
    if (B.R[k]) {
      NAS3D_REQUIRE(B.mr[k] && B.gamma[k] && B.a[k] && B.b[k] && B.p[k] && B.q[k] && B.r[k] &&
                        B.dgamma[k] && B.dbeta[k],
                    "affine_sum_bwd_apply_gn: GroupNorm term %d lacks an array", k);
      NAS3D_REQUIRE(B.dbias[k] == nullptr || B.S[k] != nullptr, "affine_sum_bwd_apply_gn: dbias needs S");
    }
  }
  const int C4 = C / 4;
  const long long per_sample = V * C4;
  NAS3D_REQUIRE(per_sample < (1ll << 31), "affine_sum_bwd_apply_gn: sample too large");
  const double inv_m = 1.0 / ((double)(C / G) * (double)V);
  affine_sum_bwd_apply_gn_kernel<<<grid2d(per_sample, N, 512), 256, 0, (cudaStream_t)stream>>>(
      T, B, G, inv_m, dout, ld_dout, V, C, C4, (unsigned)per_sample);
  return launched("affine_sum_bwd_apply_gn");
}

int nas3d_affine_sum_bwd_apply(int nterms, const float* const* x, const int* ld_x,
                               const float* const* a, const float* const* b, const int* relu,
                               const float* const* p, const float* const* q,
                               const float* const* r, const float* const* w,
                               float* const* dx, const int* ld_dx, const int* accumulate,
                               const float* dout, int ld_dout, int N, long long V, int C,
                               void* stream) {
  NAS3D_REQUIRE(nterms >= 1 && nterms <= NAS3D_MAX_TERMS, "affine_sum_bwd_apply: nterms=%d", nterms);
  NAS3D_REQUIRE(C % 4 == 0 && ld_dout % 4 == 0 && aligned16(dout),
                "affine_sum_bwd_apply: C=%d ld_dout=%d", C, ld_dout);
  BwdTerms T;
  T.nterms = nterms;
  for (int k = 0; k < nterms; ++k) {
    NAS3D_REQUIRE(ld_dx[k] % 4 == 0 && aligned16(dx[k]), "affine_sum_bwd_apply: dx %d misaligned", k);
    T.x[k] = x ? x[k] : nullptr; T.ld[k] = ld_x ? ld_x[k] : 0;
    T.a[k] = a ? a[k] : nullptr; T.b[k] = b ? b[k] : nullptr;
    T.p[k] = p ? p[k] : nullptr; T.q[k] = q ? q[k] : nullptr; T.r[k] = r ? r[k] : nullptr;
    T.w[k] = w ? w[k] : nullptr;
    T.dx[k] = dx[k]; T.ld_dx[k] = ld_dx[k];
    T.relu[k] = relu ? relu[k] : 0; T.acc[k] = accumulate ? accumulate[k] : 0;
    NAS3D_REQUIRE(!(T.relu[k] || T.q[k]) || T.x[k], "affine_sum_bwd_apply: term %d needs x", k);
  }
  const int C4 = C / 4;
  const long long per_sample = V * C4;
  NAS3D_REQUIRE(per_sample < (1ll << 31), "affine_sum_bwd_apply: sample too large");
  // ring-staged kernel (option apply_ring = 0: the register-staged kernel).  Measured on B200:
  // 426.9 -> 432.7 patches/s on the searched-net step (profiles/r2a_ab_optin_variants.txt)
  if (g_opt.apply_ring && nterms <= 2 && per_sample * N >= ring_min_elems() &&
      (C4 == 3 || (C4 & (C4 - 1)) == 0) && C4 <= 256 && (nterms < 2 || T.dx[0] != T.dx[1])) {
    int rc = nterms == 1
                 ? launch_bwd_apply_ring<1>(T, dout, ld_dout, N, V, C, C4, per_sample, (cudaStream_t)stream)
                 : launch_bwd_apply_ring<2>(T, dout, ld_dout, N, V, C, C4, per_sample, (cudaStream_t)stream);
    if (rc == NAS3D_OK) return launched("affine_sum_bwd_apply_ring");
    if (rc != NAS3D_ERR_UNSUPPORTED) return rc;
  }
  affine_sum_bwd_apply_kernel<<<grid2d(per_sample, N, 512), 256, 0, (cudaStream_t)stream>>>(
      T, dout, ld_dout, V, C, C4, (unsigned)per_sample);
  return launched("affine_sum_bwd_apply");
}

int nas3d_add_inplace(float* y, const float* x, long long n, void* stream) {
  NAS3D_REQUIRE(aligned16(y) && aligned16(x), "add_inplace: pointers must be 16B aligned");
  long long n4 = n / 4;
  add_inplace_kernel<<<grid_for(n4 > 0 ? n4 : 1, 256), 256, 0, (cudaStream_t)stream>>>(y, x, n4, n);
  return launched("add_inplace");
}

int nas3d_sigmoid_bwd(const float* prob, const float* dprob, float* dlogit, long long n,
                      void* stream) {
  sigmoid_bwd_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(prob, dprob, dlogit, n);
  return launched("sigmoid_bwd");
}

int nas3d_ncdhw_to_ndhwc(const float* src, float* dst, int N, int C, long long V, int ld_dst,
                         void* stream) {
  NAS3D_REQUIRE(ld_dst >= C, "ncdhw_to_ndhwc: ld_dst < C");
  const long long total = (long long)N * V;
  ncdhw_to_ndhwc_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, C, V,
                                                                                ld_dst, total);
  return launched("ncdhw_to_ndhwc");
}

}  // extern "C"
