// Per-(n,c) reductions and the tiny coefficient kernels that turn them into the per-(n,c)
// affine coefficients consumed by the fused affine-sum kernels (affine.cu).
//
// Memory-bound family (HBM roofline): each activation tensor is read exactly once, with
// 128-bit loads; fp32 partials over <=32 elements per thread, fp64 from there on
// (GroupNorm statistics of inputs up to ~110 in magnitude cancel badly in fp32).
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "gn_coef.cuh"

namespace nas3d {

thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};
VariantCounter g_variants[kMaxVariants];

constexpr int RB = 256;     // threads per reduction block
constexpr int RITER = 32;   // super-elements per thread

// ---------------------------------------------------------------------------------------
// moments: S[n][c] = {sum x, sum x^2}
// channel float4-groups C4 = U*P (U odd, P power of two <= 256).  A thread owns the U groups
// p*U..p*U+U-1 of the voxels j>>log2(P) for its fixed p = tid & (P-1).
// ---------------------------------------------------------------------------------------
template <int U>
__global__ void __launch_bounds__(RB) moments_nc_kernel(const float* __restrict__ x, long long V,
                                                         int C, int ld, int P, int logP,
                                                         double* __restrict__ S, int iters) {
  const int n = blockIdx.y;
  const long long total = V << logP;  // super-elements in this sample
  const long long j0 = (long long)blockIdx.x * (RB * iters) + threadIdx.x;
  const int p = threadIdx.x & (P - 1);
  const float* xb = x + (long long)n * V * ld + (long long)p * U * 4;

  float s[U][4], q[U][4];
#pragma unroll
  for (int u = 0; u < U; ++u)
#pragma unroll
    for (int e = 0; e < 4; ++e) s[u][e] = q[u][e] = 0.f;

#pragma unroll 4
  for (int it = 0; it < iters; ++it) {
    long long j = j0 + (long long)it * RB;
    if (j < total) {
      const float* px = xb + (j >> logP) * ld;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float4 v = ldg4(px + u * 4);
        s[u][0] += v.x; s[u][1] += v.y; s[u][2] += v.z; s[u][3] += v.w;
        q[u][0] += v.x * v.x; q[u][1] += v.y * v.y; q[u][2] += v.z * v.z; q[u][3] += v.w * v.w;
      }
    }
  }

  __shared__ double sm[2 * 4 * 64 * 3 / 1];  // [C*2] (C <= 768)
  for (int i = threadIdx.x; i < C * 2; i += RB) sm[i] = 0.0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int u = 0; u < U; ++u)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      double a = s[u][e], b = q[u][e];
      for (int o = 16; o >= P; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      if (lane < P || P > 32) {
        int c = (p * U + u) * 4 + e;
        atomicAdd(&sm[c * 2 + 0], a);
        atomicAdd(&sm[c * 2 + 1], b);
      }
    }
  __syncthreads();
  double* Sn = S + (long long)n * C * 2;
  for (int i = threadIdx.x; i < C * 2; i += RB) atomicAdd(&Sn[i], sm[i]);
}

// ---------------------------------------------------------------------------------------
// backward reduce for up to TG terms per block (blockIdx.z selects the term group):
//   R[k][n][c] = { sum m*dout, sum m*dout*x_k },  m = (a*x+b > 0) if relu else 1
// ---------------------------------------------------------------------------------------
struct ReduceTerms {
  const float* x[NAS3D_MAX_TERMS];
  const float* a[NAS3D_MAX_TERMS];
  const float* b[NAS3D_MAX_TERMS];
  double* R[NAS3D_MAX_TERMS];
  int ld[NAS3D_MAX_TERMS];
  int relu[NAS3D_MAX_TERMS];
  int nterms;
};

template <int U, int TG>
__global__ void __launch_bounds__(RB)
    bwd_reduce_kernel(const __grid_constant__ ReduceTerms T, const float* __restrict__ dout,
                      int ld_dout, long long V, int C, int P, int logP, int iters) {
  const int n = blockIdx.y;
  const int k0 = blockIdx.z * TG;
  const long long total = V << logP;
  const long long j0 = (long long)blockIdx.x * (RB * iters) + threadIdx.x;
  const int p = threadIdx.x & (P - 1);
  const int cbase = p * U * 4;
  const float* db = dout + (long long)n * V * ld_dout + cbase;

  float r1[TG][U][4], r2[TG][U][4];
#pragma unroll
  for (int t = 0; t < TG; ++t)
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int e = 0; e < 4; ++e) r1[t][u][e] = r2[t][u][e] = 0.f;

  // Two voxels per thread per trip, all their dout / x_k loads issued before the first use
  // (predicated, no early exit in the body): twice the bytes in flight per thread.
  const float* xb[TG];
  float4 a4[TG][U], b4[TG][U];
  bool live[TG], relu[TG];
#pragma unroll
  for (int t = 0; t < TG; ++t) {
    const int k = k0 + t;
    live[t] = k < T.nterms;
    relu[t] = live[t] && T.relu[k];
    xb[t] = live[t] ? T.x[k] + (long long)n * V * T.ld[k] + cbase : nullptr;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      a4[t][u] = make_float4(1.f, 1.f, 1.f, 1.f);
      b4[t][u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (relu[t] && T.a[k]) a4[t][u] = ldg4(T.a[k] + (long long)n * C + cbase + u * 4);
      if (relu[t] && T.b[k]) b4[t][u] = ldg4(T.b[k] + (long long)n * C + cbase + u * 4);
    }
  }
#pragma unroll 1
  for (int it = 0; it < iters; it += 2) {
    const long long ja = j0 + (long long)it * RB, jb = ja + RB;
    const bool va = ja < total, vb = (it + 1 < iters) && jb < total;
    if (!va) break;
    const long long voxa = ja >> logP, voxb = vb ? (jb >> logP) : voxa;
    float4 da[U], dbv[U], xa[TG][U], xbv[TG][U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      da[u] = ldg4(db + voxa * ld_dout + u * 4);
      dbv[u] = ldg4(db + voxb * ld_dout + u * 4);
#pragma unroll
      for (int t = 0; t < TG; ++t) {
        xa[t][u] = xbv[t][u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live[t]) {
          const int ldk = T.ld[k0 + t];
          xa[t][u] = ldg4(xb[t] + voxa * ldk + u * 4);
          xbv[t][u] = ldg4(xb[t] + voxb * ldk + u * 4);
        }
      }
    }
    const float wb = vb ? 1.f : 0.f;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float d0[4] = {da[u].x, da[u].y, da[u].z, da[u].w};
      const float d1[4] = {dbv[u].x * wb, dbv[u].y * wb, dbv[u].z * wb, dbv[u].w * wb};
#pragma unroll
      for (int t = 0; t < TG; ++t) {
        if (!live[t]) continue;
        const float x0[4] = {xa[t][u].x, xa[t][u].y, xa[t][u].z, xa[t][u].w};
        const float x1[4] = {xbv[t][u].x, xbv[t][u].y, xbv[t][u].z, xbv[t][u].w};
        const float av[4] = {a4[t][u].x, a4[t][u].y, a4[t][u].z, a4[t][u].w};
        const float bv[4] = {b4[t][u].x, b4[t][u].y, b4[t][u].z, b4[t][u].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float m0 = (!relu[t] || av[e] * x0[e] + bv[e] > 0.f) ? d0[e] : 0.f;
          const float m1 = (!relu[t] || av[e] * x1[e] + bv[e] > 0.f) ? d1[e] : 0.f;
          r1[t][u][e] += m0;
          r2[t][u][e] += m0 * x0[e];
          r1[t][u][e] += m1;
          r2[t][u][e] += m1 * x1[e];
        }
      }
    }
  }

  __shared__ double sm[TG * 768 * 2];  // TG * C * 2, C <= 768 (TG=4 -> 48 KB is too much: see host)
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < TG * C * 2; i += RB) sm[i] = 0.0;
  __syncthreads();
#pragma unroll
  for (int t = 0; t < TG; ++t)
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        double a = r1[t][u][e], b = r2[t][u][e];
        for (int o = 16; o >= P; o >>= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, o);
          b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if (lane < P || P > 32) {
          int c = cbase + u * 4 + e;
          atomicAdd(&sm[(t * C + c) * 2 + 0], a);
          atomicAdd(&sm[(t * C + c) * 2 + 1], b);
        }
      }
  __syncthreads();
  for (int i = threadIdx.x; i < TG * C * 2; i += RB) {
    int t = i / (C * 2);
    int k = k0 + t;
    if (k < T.nterms) atomicAdd(&T.R[k][(long long)n * C * 2 + (i - t * C * 2)], sm[i]);
  }
}

// ---------------------------------------------------------------------------------------
// Ring-staged variant of bwd_reduce for the big narrow tensors (C = 4*P <= 64, P a power of two):
// same mapping, sums and epilogue, but the dout / x_k float4s of a thread's next RR_S-1 trips are
// in flight as cp.async copies into the thread's own slots of a shared-memory ring (as in
// conv_pointwise_bwd.cu) instead of being held in registers: this read-only kernel ran at
// 4.3-4.6 TB/s with two trips in flight per thread (ncu: long_scoreboard), while the U = 3
// instantiation, which has 2x the bytes in flight per thread, reaches 6.3 TB/s.
// ---------------------------------------------------------------------------------------
constexpr int RR_S = 5;

__device__ __forceinline__ void rr_cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(
                   (unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

template <int TG>
__global__ void __launch_bounds__(RB)
    bwd_reduce_ring_kernel(const __grid_constant__ ReduceTerms T, const float* __restrict__ dout,
                           int ld_dout, long long V, int C, int P, int logP, int iters) {
  extern __shared__ __align__(16) float4 rring[];      // [component][stage][thread]
  __shared__ double sm[TG * 64 * 2];                   // C <= 64
  const int n = blockIdx.y;
  const int k0 = blockIdx.z * TG;
  const long long total = V << logP;
  const long long j0 = (long long)blockIdx.x * (RB * iters) + threadIdx.x;
  const int p = threadIdx.x & (P - 1);
  const int cbase = p * 4;
  const int my_iters = j0 < total ? (int)min((long long)iters, (total - j0 + RB - 1) / RB) : 0;
  const long long vox0 = j0 >> logP;                   // trip i reads voxel vox0 + i * (RB >> logP)
  const long long vstep = RB >> logP;

  bool live[TG], relu[TG];
  float4 a4[TG], b4[TG];
  const float* px[TG];
  long long sx[TG];
#pragma unroll
  for (int t = 0; t < TG; ++t) {
    const int k = k0 + t;
    live[t] = k < T.nterms;
    relu[t] = live[t] && T.relu[k];
    a4[t] = make_float4(1.f, 1.f, 1.f, 1.f);
    b4[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    px[t] = nullptr; sx[t] = 0;
    if (live[t]) {
      px[t] = T.x[k] + ((long long)n * V + vox0) * T.ld[k] + cbase;
      sx[t] = vstep * T.ld[k];
      if (relu[t] && T.a[k]) a4[t] = ldg4(T.a[k] + (long long)n * C + cbase);
      if (relu[t] && T.b[k]) b4[t] = ldg4(T.b[k] + (long long)n * C + cbase);
    }
  }
  const float* pd = dout + ((long long)n * V + vox0) * ld_dout + cbase;
  const long long sd = vstep * ld_dout;

  float4* const r0 = rring + threadIdx.x;
  float4* const r_end = r0 + RR_S * RB;
  float4* is = r0;
  int issued = 0;
  auto issue = [&]() {
    if (issued < my_iters) {
      rr_cp16(is, pd);
#pragma unroll
      for (int t = 0; t < TG; ++t)
        if (live[t]) rr_cp16(is + (1 + t) * RR_S * RB, px[t]);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    ++issued;
    pd += sd;
#pragma unroll
    for (int t = 0; t < TG; ++t) px[t] += sx[t];
    is += RB;
    if (is == r_end) is = r0;
  };
#pragma unroll
  for (int s = 0; s < RR_S - 1; ++s) issue();

  float r1[TG][4], r2[TG][4];
#pragma unroll
  for (int t = 0; t < TG; ++t)
#pragma unroll
    for (int e = 0; e < 4; ++e) r1[t][e] = r2[t][e] = 0.f;

  const float4* cs = r0;
#pragma unroll 1
  for (int it = 0; it < my_iters; ++it) {
    issue();
    asm volatile("cp.async.wait_group %0;\n" ::"n"(RR_S - 1) : "memory");
    const float4 d4 = cs[0];
    const float d[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
    for (int t = 0; t < TG; ++t) {
      if (!live[t]) continue;
      const float4 x4 = cs[(1 + t) * RR_S * RB];
      const float x[4] = {x4.x, x4.y, x4.z, x4.w};
      const float av[4] = {a4[t].x, a4[t].y, a4[t].z, a4[t].w};
      const float bv[4] = {b4[t].x, b4[t].y, b4[t].z, b4[t].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float m = (!relu[t] || av[e] * x[e] + bv[e] > 0.f) ? d[e] : 0.f;
        r1[t][e] += m;
        r2[t][e] += m * x[e];
      }
    }
    cs += RB;
    if (cs == r_end) cs = r0;
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");

  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < TG * C * 2; i += RB) sm[i] = 0.0;
  __syncthreads();
#pragma unroll
  for (int t = 0; t < TG; ++t)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      double a = r1[t][e], b = r2[t][e];
      for (int o = 16; o >= P; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      if (lane < P || P > 32) {
        const int c = cbase + e;
        atomicAdd(&sm[(t * C + c) * 2 + 0], a);
        atomicAdd(&sm[(t * C + c) * 2 + 1], b);
      }
    }
  __syncthreads();
  for (int i = threadIdx.x; i < TG * C * 2; i += RB) {
    const int t = i / (C * 2);
    const int k = k0 + t;
    if (k < T.nterms) atomicAdd(&T.R[k][(long long)n * C * 2 + (i - t * C * 2)], sm[i]);
  }
}

// ---------------------------------------------------------------------------------------
// tiny per-sample coefficient kernels: grid = N, block = 128
// ---------------------------------------------------------------------------------------
constexpr int TB = 128;

__device__ __forceinline__ double block_sum(double v, double* sh) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < TB / 32; ++i) t += sh[i];
  return t;
}

__global__ void __launch_bounds__(TB)
    gn_coef_kernel(const double* __restrict__ S, const float* __restrict__ gamma,
                   const float* __restrict__ beta, int C, int G, double inv_m, float eps,
                   float* __restrict__ a, float* __restrict__ b, float* __restrict__ mean_rstd) {
  const int n = blockIdx.x;
  const int cg = C / G;
  __shared__ float sh_mean[64], sh_rstd[64];
  const double* Sn = S + (long long)n * C * 2;
  for (int g = threadIdx.x; g < G; g += TB) {
    double s = 0.0, q = 0.0;
    for (int c = g * cg; c < (g + 1) * cg; ++c) { s += Sn[c * 2]; q += Sn[c * 2 + 1]; }
    double mean = s * inv_m;
    double var = q * inv_m - mean * mean;
    if (var < 0.0) var = 0.0;
    float rstd = (float)(1.0 / sqrt(var + (double)eps));
    sh_mean[g] = (float)mean;
    sh_rstd[g] = rstd;
    mean_rstd[((long long)n * G + g) * 2 + 0] = (float)mean;
    mean_rstd[((long long)n * G + g) * 2 + 1] = rstd;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += TB) {
    int g = c / cg;
    float av = sh_rstd[g] * gamma[c];
    a[(long long)n * C + c] = av;
    b[(long long)n * C + c] = beta[c] - sh_mean[g] * av;
  }
}

__global__ void __launch_bounds__(TB)
    se_excite_kernel(const double* __restrict__ S, const float* __restrict__ W1,
                     const float* __restrict__ b1, const float* __restrict__ W2,
                     const float* __restrict__ b2, int C, double inv_v, float* __restrict__ s,
                     float* __restrict__ hz) {
  const int n = blockIdx.x;
  __shared__ double sh[TB / 32];
  double part = 0.0;
  for (int c = threadIdx.x; c < C; c += TB)
    part += (double)W1[c] * (double)(float)(S[((long long)n * C + c) * 2] * inv_v);
  float z = (float)block_sum(part, sh) + b1[0];
  float h = z > 0.f ? z : 0.f;
  if (threadIdx.x == 0) { hz[n * 2 + 0] = h; hz[n * 2 + 1] = z; }
  for (int c = threadIdx.x; c < C; c += TB) {
    float e = W2[c] * h + b2[c];
    s[(long long)n * C + c] = 1.f / (1.f + expf(-e));
  }
}

__global__ void __launch_bounds__(TB)
    gn_bwd_coef_kernel(const double* __restrict__ R, const float* __restrict__ mean_rstd,
                       const float* __restrict__ gamma, const float* __restrict__ a,
                       const float* __restrict__ b, const float* __restrict__ w, int C, int G,
                       double inv_m, float* __restrict__ p, float* __restrict__ q,
                       float* __restrict__ r, float* __restrict__ dgamma,
                       float* __restrict__ dbeta, float* __restrict__ dw,
                       const double* __restrict__ S, double V, float* __restrict__ dbias) {
  const int n = blockIdx.x;
  const int cg = C / G;
  const double wv = w ? (double)w[0] : 1.0;
  __shared__ double shA[64], shB[64];
  __shared__ double sh[TB / 32];
  const double* Rn = R + (long long)n * C * 2;
  for (int g = threadIdx.x; g < G; g += TB) {
    double mu = mean_rstd[((long long)n * G + g) * 2 + 0];
    double rho = mean_rstd[((long long)n * G + g) * 2 + 1];
    double A = 0.0, B = 0.0;
    for (int c = g * cg; c < (g + 1) * cg; ++c) {
      double r1 = Rn[c * 2], r2 = Rn[c * 2 + 1];
      A += (double)gamma[c] * r1;
      B += (double)gamma[c] * rho * (r2 - mu * r1);
    }
    shA[g] = A;
    shB[g] = B;
  }
  __syncthreads();
  double dwp = 0.0;
  for (int c = threadIdx.x; c < C; c += TB) {
    int g = c / cg;
    double mu = mean_rstd[((long long)n * G + g) * 2 + 0];
    double rho = mean_rstd[((long long)n * G + g) * 2 + 1];
    double r1 = Rn[c * 2], r2 = Rn[c * 2 + 1];
    double qq = -rho * rho * shB[g] * inv_m;
    p[(long long)n * C + c] = (float)(wv * rho * (double)gamma[c]);
    q[(long long)n * C + c] = (float)(wv * qq);
    r[(long long)n * C + c] = (float)(wv * (-qq * mu - rho * shA[g] * inv_m));
    atomicAdd(&dgamma[c], (float)(wv * rho * (r2 - mu * r1)));
    atomicAdd(&dbeta[c], (float)(wv * r1));
    if (dbias) {
      // bias gradient of the conv that produced x:  sum_v dx = p*R1 + q*sum(x) + r*V  (analytic,
      // saves a pass over the gradient tensor)
      const double pp = wv * rho * (double)gamma[c];
      const double rr = wv * (-qq * mu - rho * shA[g] * inv_m);
      atomicAdd(&dbias[c], (float)(pp * r1 + wv * qq * S[((long long)n * C + c) * 2] + rr * V));
    }
    dwp += (double)a[(long long)n * C + c] * r2 + (double)b[(long long)n * C + c] * r1;
  }
  if (dw) {
    double tot = block_sum(dwp, sh);
    if (threadIdx.x == 0) atomicAdd(dw, (float)tot);
  }
}

__global__ void __launch_bounds__(TB)
    se_bwd_coef_kernel(const double* __restrict__ R, const double* __restrict__ S,
                       const float* __restrict__ s, const float* __restrict__ hz,
                       const float* __restrict__ W1, const float* __restrict__ W2,
                       const float* __restrict__ w, int C, double inv_v, float* __restrict__ p,
                       float* __restrict__ r, float* __restrict__ dW1, float* __restrict__ db1,
                       float* __restrict__ dW2, float* __restrict__ db2, float* __restrict__ dw) {
  const int n = blockIdx.x;
  const double wv = w ? (double)w[0] : 1.0;
  const float h = hz[n * 2 + 0], z = hz[n * 2 + 1];
  __shared__ double sh[TB / 32];
  double dh_part = 0.0, dw_part = 0.0;
  for (int c = threadIdx.x; c < C; c += TB) {
    double r2 = R[((long long)n * C + c) * 2 + 1];
    double sv = s[(long long)n * C + c];
    double de = wv * r2 * sv * (1.0 - sv);
    atomicAdd(&dW2[c], (float)(de * h));
    atomicAdd(&db2[c], (float)de);
    dh_part += de * (double)W2[c];
    dw_part += sv * r2;
    p[(long long)n * C + c] = (float)(wv * sv);
  }
  double dh = block_sum(dh_part, sh);
  double dz = z > 0.f ? dh : 0.0;
  for (int c = threadIdx.x; c < C; c += TB) {
    double mean = S[((long long)n * C + c) * 2] * inv_v;
    atomicAdd(&dW1[c], (float)(dz * mean));
    r[(long long)n * C + c] = (float)(dz * (double)W1[c] * inv_v);
  }
  if (threadIdx.x == 0) atomicAdd(db1, (float)dz);
  if (dw) {
    double tot = block_sum(dw_part, sh);
    if (threadIdx.x == 0) atomicAdd(dw, (float)tot);
  }
}

__global__ void __launch_bounds__(TB)
    plain_bwd_coef_kernel(const double* __restrict__ R, int NC, float* __restrict__ dw) {
  __shared__ double sh[TB / 32];
  double part = 0.0;
  for (int i = threadIdx.x; i < NC; i += TB) part += R[(long long)i * 2 + 1];
  double tot = block_sum(part, sh);
  if (threadIdx.x == 0) atomicAdd(dw, (float)tot);
}

// ---- batched variants: all GroupNorm terms of one node in a single launch (grid = N x nterms);
// the supernet has up to 22 terms per node and these tiny kernels are pure launch latency.  The
// affine kernels can also run the same bodies in their own prologue (affine.cu, *_gn entry points).
__global__ void __launch_bounds__(TB)
    gn_coef_batch_kernel(const __grid_constant__ GnFwdBatch B, int C, int G, double inv_m, float eps) {
  __shared__ GnScratch sc;
  gn_coef_body(B, blockIdx.y, blockIdx.x, C, G, inv_m, eps, sc);
}

__global__ void __launch_bounds__(TB)
    gn_bwd_coef_batch_kernel(const __grid_constant__ GnBwdBatch B, int C, int G, double inv_m,
                             double V) {
  __shared__ GnScratch sc;
  gn_bwd_coef_body(B, blockIdx.y, blockIdx.x, C, G, inv_m, V, true, sc);
}

// super-elements per thread: RITER for big tensors (fewest atomics), fewer when the tensor is small
// so that at least ~4 CTAs per SM share the work (deep U-Net levels are latency-, not HBM-bound)
// `resident` (CTAs of this kernel that fit on the GPU at once, 0 = unknown): a grid of 4.6 waves
// runs as long as one of 5, so the per-thread trip count is stretched to make the grid a whole
// number of waves whenever that is predicted to be faster (bwd_reduce at 8 x 128^3: 2048 CTAs of
// 32 trips = 4.61 waves -> 1776 CTAs of 37 trips = 4 waves, 4.3 -> 4.6 TB/s; at 8 x 64^3 x 8 ch:
// 608 CTAs of 27 trips = 1.37 waves -> 440 CTAs of 38 trips = 1 wave).
static int reduce_iters(long long total_per_sample, int N, int zdim, int resident = 0) {
  const long long want_ctas = (long long)kNumSMs * 4;
  long long it = total_per_sample * N * zdim / (want_ctas * RB);
  if (it < 1) it = 1;
  if (it > RITER) it = RITER;
  const bool whole_waves = g_opt.reduce_waves != 0;
  if (whole_waves && resident > 0) {
    const long long nz = (long long)N * zdim;
    const long long ctas = ((total_per_sample + RB * it - 1) / (RB * it)) * nz;
    const long long w = ctas / resident;                   // whole waves
    if (w >= 1 && ctas % resident != 0) {
      const long long gx_max = w * resident / nz;          // grid.x that fits w whole waves
      if (gx_max >= 1) {
        // w waves of it2 trips against w+1 waves of `it` trips
        const long long it2 = (total_per_sample + RB * gx_max - 1) / (RB * gx_max);
        if (it2 > it && it2 * w < it * (w + 1) && it2 <= 2 * RITER) it = it2;
      }
    }
  }
  return (int)it;
}

template <typename K>
static int resident_ctas(K kernel, int block) {
  int o = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kernel, block, 0) != cudaSuccess || o < 1) return 0;
  return o * kNumSMs;
}

static int check_channels(int C, int ld, int* U, int* P, int* logP) {
  NAS3D_REQUIRE(C > 0 && C % 4 == 0 && ld % 4 == 0 && ld >= C,
                "channel count %d / pitch %d must be multiples of 4", C, ld);
  NAS3D_REQUIRE(C <= 768, "C=%d > 768 unsupported", C);
  split_c4(C / 4, U, P);
  NAS3D_REQUIRE(*U == 1 || *U == 3, "C=%d: odd factor %d of C/4 unsupported (1 or 3)", C, *U);
  NAS3D_REQUIRE(*P <= 256, "C=%d too wide", C);
  int l = 0;
  while ((1 << l) < *P) ++l;
  *logP = l;
  return NAS3D_OK;
}

}  // namespace nas3d

using namespace nas3d;

extern "C" {

int nas3d_version(void) { return 100; }
const char* nas3d_last_error(void) { return g_err; }
unsigned long long nas3d_launch_count(void) { return g_launches.load(); }
unsigned long long nas3d_launch_count_of(const char* label) {
  unsigned long long n = 0;
  if (!label) return 0;
  for (int i = 0; i < kMaxVariants; ++i) {
    const char* cur = g_variants[i].name.load(std::memory_order_relaxed);
    if (cur && strcmp(cur, label) == 0) n += g_variants[i].n.load(std::memory_order_relaxed);
  }
  return n;
}
int nas3d_launch_labels(char* buf, int cap) {
  int used = 0;
  if (buf && cap > 0) buf[0] = 0;
  for (int i = 0; i < kMaxVariants; ++i) {
    const char* cur = g_variants[i].name.load(std::memory_order_relaxed);
    if (!cur) continue;
    const int len = (int)strlen(cur);
    if (buf && used + len + 2 <= cap) {
      memcpy(buf + used, cur, len);
      buf[used + len] = '\n';
      buf[used + len + 1] = 0;
    }
    used += len + 1;
  }
  return used + 1;
}

int nas3d_moments_nc(const float* x, int N, long long V, int C, int ld, double* S, void* stream) {
  int U, P, logP;
  int rc = check_channels(C, ld, &U, &P, &logP);
  if (rc) return rc;
  NAS3D_REQUIRE(aligned16(x), "moments_nc: x not 16B aligned");
  cudaStream_t st = (cudaStream_t)stream;
  NAS3D_CUDA(cudaMemsetAsync(S, 0, sizeof(double) * 2 * (size_t)N * C, st));
  long long total = V * P;
  const int iters = reduce_iters(total, N, 1);
  dim3 grid((unsigned)((total + RB * iters - 1) / (RB * iters)), N);
  if (U == 1) moments_nc_kernel<1><<<grid, RB, 0, st>>>(x, V, C, ld, P, logP, S, iters);
  else moments_nc_kernel<3><<<grid, RB, 0, st>>>(x, V, C, ld, P, logP, S, iters);
  return launched("moments_nc");
}

int nas3d_gn_coef(const double* S, const float* gamma, const float* beta, int N, int C, int G,
                  long long V, float eps, float* a, float* b, float* mean_rstd, void* stream) {
  NAS3D_REQUIRE(G >= 1 && G <= 64 && C % G == 0, "gn_coef: bad groups %d for C=%d", G, C);
  double inv_m = 1.0 / ((double)(C / G) * (double)V);
  gn_coef_kernel<<<N, TB, 0, (cudaStream_t)stream>>>(S, gamma, beta, C, G, inv_m, eps, a, b,
                                                     mean_rstd);
  return launched("gn_coef");
}

int nas3d_gn_coef_batch(int nterms, const double* const* S, const float* const* gamma,
                        const float* const* beta, int N, int C, int G, long long V, float eps,
                        float* const* a, float* const* b, float* const* mean_rstd, void* stream) {
  NAS3D_REQUIRE(nterms >= 1 && nterms <= NAS3D_MAX_TERMS, "gn_coef_batch: nterms=%d", nterms);
  NAS3D_REQUIRE(G >= 1 && G <= 64 && C % G == 0, "gn_coef_batch: bad groups %d for C=%d", G, C);
  GnFwdBatch B;
  for (int k = 0; k < nterms; ++k) {
    B.S[k] = S[k]; B.gamma[k] = gamma[k]; B.beta[k] = beta[k];
    B.a[k] = a[k]; B.b[k] = b[k]; B.mr[k] = mean_rstd[k];
  }
  const double inv_m = 1.0 / ((double)(C / G) * (double)V);
  gn_coef_batch_kernel<<<dim3(N, nterms), TB, 0, (cudaStream_t)stream>>>(B, C, G, inv_m, eps);
  return launched("gn_coef_batch");
}

int nas3d_gn_bwd_coef_batch(int nterms, const double* const* R, const float* const* mean_rstd,
                            const float* const* gamma, const float* const* a,
                            const float* const* b, const float* const* w, int N, int C, int G,
                            long long V, float* const* p, float* const* q, float* const* r,
                            float* const* dgamma, float* const* dbeta, float* const* dw,
                            const double* const* S, float* const* dbias_prev, void* stream) {
  NAS3D_REQUIRE(nterms >= 1 && nterms <= NAS3D_MAX_TERMS, "gn_bwd_coef_batch: nterms=%d", nterms);
  NAS3D_REQUIRE(G >= 1 && G <= 64 && C % G == 0, "gn_bwd_coef_batch: bad groups %d for C=%d", G, C);
  GnBwdBatch B;
  for (int k = 0; k < nterms; ++k) {
    B.R[k] = R[k]; B.mr[k] = mean_rstd[k]; B.gamma[k] = gamma[k]; B.a[k] = a[k]; B.b[k] = b[k];
    B.w[k] = w ? w[k] : nullptr; B.p[k] = p[k]; B.q[k] = q[k]; B.r[k] = r[k];
    B.dgamma[k] = dgamma[k]; B.dbeta[k] = dbeta[k]; B.dw[k] = dw ? dw[k] : nullptr;
    B.S[k] = S ? S[k] : nullptr; B.dbias[k] = dbias_prev ? dbias_prev[k] : nullptr;
    NAS3D_REQUIRE(B.dbias[k] == nullptr || B.S[k] != nullptr, "gn_bwd_coef_batch: dbias needs S");
  }
  const double inv_m = 1.0 / ((double)(C / G) * (double)V);
  gn_bwd_coef_batch_kernel<<<dim3(N, nterms), TB, 0, (cudaStream_t)stream>>>(B, C, G, inv_m, (double)V);
  return launched("gn_bwd_coef_batch");
}

int nas3d_se_excite(const double* S, const float* W1, const float* b1, const float* W2,
                    const float* b2, int N, int C, long long V, float* s, float* hz, void* stream) {
  se_excite_kernel<<<N, TB, 0, (cudaStream_t)stream>>>(S, W1, b1, W2, b2, C, 1.0 / (double)V, s,
                                                       hz);
  return launched("se_excite");
}

int nas3d_affine_sum_bwd_reduce(int nterms, const float* const* x, const int* ld_x,
                                const float* const* a, const float* const* b, const int* relu,
                                const float* dout, int ld_dout, double* const* R, int N,
                                long long V, int C, void* stream) {
  NAS3D_REQUIRE(nterms >= 1 && nterms <= NAS3D_MAX_TERMS, "bwd_reduce: nterms=%d", nterms);
  int U, P, logP;
  int rc = check_channels(C, ld_dout, &U, &P, &logP);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  ReduceTerms T;
  T.nterms = nterms;
  // R buffers carved from one allocation are cleared with a single memset node
  const size_t rbytes = sizeof(double) * 2 * (size_t)N * C;
  bool contiguous = true;
  for (int k = 0; k < nterms; ++k) {
    NAS3D_REQUIRE(ld_x[k] % 4 == 0 && aligned16(x[k]), "bwd_reduce: term %d misaligned", k);
    T.x[k] = x[k]; T.a[k] = a ? a[k] : nullptr; T.b[k] = b ? b[k] : nullptr;
    T.R[k] = R[k]; T.ld[k] = ld_x[k]; T.relu[k] = relu ? relu[k] : 0;
    if (k > 0 && (const char*)R[k] != (const char*)R[k - 1] + rbytes) contiguous = false;
  }
  if (contiguous) {
    NAS3D_CUDA(cudaMemsetAsync(R[0], 0, rbytes * nterms, st));
  } else {
    for (int k = 0; k < nterms; ++k) NAS3D_CUDA(cudaMemsetAsync(R[k], 0, rbytes, st));
  }
  long long total = V * P;
  // TG=2 keeps the fp64 staging array (TG*C*2 doubles) inside 48 KB static smem for C<=768/..;
  // wide tensors are tiny in this network so TG=1 there.
  // option reduce_ring = 0: the register-staged kernel.
  // Measured on B200: affine_sum_bwd_reduce 2.14 -> 1.79 ms per searched-net step, 416.6 -> 420.7
  // patches/s (profiles/r1f_ab_reduce_ring.json)
  const bool use_ring = g_opt.reduce_ring != 0;
  if (use_ring && U == 1 && C <= 64 && total * N >= ring_min_elems()) {
    auto kern = bwd_reduce_ring_kernel<2>;
    const int smem = 3 * RR_S * RB * (int)sizeof(float4);
    static int resident = -1;
    if (resident < 0) {
      NAS3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      int o = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, RB, smem) != cudaSuccess || o < 1) o = 0;
      resident = o * kNumSMs;
    }
    const int zd = (nterms + 1) / 2, iters = reduce_iters(total, N, zd, resident);
    dim3 grid((unsigned)((total + RB * iters - 1) / (RB * iters)), N, zd);
    kern<<<grid, RB, smem, st>>>(T, dout, ld_dout, V, C, P, logP, iters);
    return launched("affine_sum_bwd_reduce_ring");
  }
  if (U == 1 && C <= 256) {
    static const int resident = resident_ctas(bwd_reduce_kernel<1, 2>, RB);
    const int zd = (nterms + 1) / 2, iters = reduce_iters(total, N, zd, resident);
    dim3 grid((unsigned)((total + RB * iters - 1) / (RB * iters)), N, zd);
    bwd_reduce_kernel<1, 2><<<grid, RB, 0, st>>>(T, dout, ld_dout, V, C, P, logP, iters);
  } else if (U == 1) {
    const int iters = reduce_iters(total, N, nterms);
    dim3 grid((unsigned)((total + RB * iters - 1) / (RB * iters)), N, nterms);
    bwd_reduce_kernel<1, 1><<<grid, RB, 0, st>>>(T, dout, ld_dout, V, C, P, logP, iters);
  } else {
    const int iters = reduce_iters(total, N, nterms);
    dim3 grid((unsigned)((total + RB * iters - 1) / (RB * iters)), N, nterms);
    bwd_reduce_kernel<3, 1><<<grid, RB, 0, st>>>(T, dout, ld_dout, V, C, P, logP, iters);
  }
  return launched("affine_sum_bwd_reduce");
}

int nas3d_gn_bwd_coef(const double* R, const float* mean_rstd, const float* gamma,
                      const float* a, const float* b, const float* w, int N, int C, int G,
                      long long V, float* p, float* q, float* r, float* dgamma, float* dbeta,
                      float* dw, const double* S, float* dbias_prev, void* stream) {
  NAS3D_REQUIRE(G >= 1 && G <= 64 && C % G == 0, "gn_bwd_coef: bad groups %d for C=%d", G, C);
  NAS3D_REQUIRE(dbias_prev == nullptr || S != nullptr, "gn_bwd_coef: dbias_prev needs the moments S");
  double inv_m = 1.0 / ((double)(C / G) * (double)V);
  gn_bwd_coef_kernel<<<N, TB, 0, (cudaStream_t)stream>>>(R, mean_rstd, gamma, a, b, w, C, G,
                                                         inv_m, p, q, r, dgamma, dbeta, dw, S,
                                                         (double)V, dbias_prev);
  return launched("gn_bwd_coef");
}

int nas3d_se_bwd_coef(const double* R, const double* S, const float* s, const float* hz,
                      const float* W1, const float* W2, const float* w, int N, int C, long long V,
                      float* p, float* r, float* dW1, float* db1, float* dW2, float* db2,
                      float* dw, void* stream) {
  se_bwd_coef_kernel<<<N, TB, 0, (cudaStream_t)stream>>>(R, S, s, hz, W1, W2, w, C,
                                                         1.0 / (double)V, p, r, dW1, db1, dW2,
                                                         db2, dw);
  return launched("se_bwd_coef");
}

int nas3d_plain_bwd_coef(const double* R, const float* w, int N, int C, float* dw, void* stream) {
  (void)w;
  NAS3D_REQUIRE(dw != nullptr, "plain_bwd_coef: dw is NULL");
  plain_bwd_coef_kernel<<<1, TB, 0, (cudaStream_t)stream>>>(R, N * C, dw);
  return launched("plain_bwd_coef");
}

}  // extern "C"
