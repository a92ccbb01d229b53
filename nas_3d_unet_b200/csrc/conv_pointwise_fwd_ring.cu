// Ring-staged forward of the big stride-1 1x1x1 convolutions (stems, cell preprocess convs over a
// dense tensor or a virtual concat, the 12->3 head):  small[v,cs] = act(bias + sum_cb f(big[v,cb]) W[cs][cb]).
//
// Validated and on by default since round 2 (library option pw_fwd_ring; parity test
// test_ring_staged_pointwise_forward_matches_default_kernel).  Motivation: the register-staged
// pointwise_kernel runs these shapes at 3.3-5.1 TB/s (ncu: long_scoreboard at 18-25 % occupancy,
// 96-162 registers), while the same per-thread cp.async ring took bwd_reduce / affine_sum_fwd / the
// fused 1x1 backward to 5.2-6.3 TB/s.
//
// Outputs wider than one float4 per voxel (CO = 8, 12) written straight from a thread-per-voxel
// mapping make every 128-bit store instruction of a warp touch all the 32-byte sectors of its
// 32 x CO x 4-byte region half-filled (the 4 -> 12 stem ran at 2.9 TB/s); for a dense destination the
// warp therefore transposes its 32 x CO outputs through a private shared-memory patch and stores
// CO/4 fully contiguous 512-byte rows instead.
//
// One thread per voxel, a CTA walks a contiguous range of PR_T-voxel tiles; the Q = Cin/4 float4
// channel groups of a thread's next PR_S-1 voxels are in flight as cp.async copies into the
// thread's own ring slots; weights sit in shared memory ([cin][CO], broadcast reads); the CO <= 12
// outputs of a voxel live in registers.  Optional fused GroupNorm statistics exactly as
// pointwise_kernel: fp32 per-thread partials over <= 32 voxels, warp sum, fp64 per-CTA array, one
// fp64 atomic per (channel, moment) per CTA and sample (host guarantees Vs % PR_T == 0).
#include <stdlib.h>
#include "common.cuh"
#include "conv_tiled.h"

namespace nas3d {

struct PwRingArgs {
  const float* src[4];     // parts of the big tensor (nseg = 1: the tensor itself)
  int ld_src[4];
  int nseg, seg_w;
  const float* w;          // W[Cout][Cin]
  const float* bias;       // optional [Cout]
  const float* scale;      // optional [N][Cin]
  float* dst;
  double* moments;         // optional [N][Cout][2]
  int Cin, Cout, ld_dst;
  unsigned nvox, Vs;
  int relu, sigmoid;
};

constexpr int PR_T = 128;
constexpr int PR_S = 4;
constexpr int PR_FLUSH = 32;     // voxels per fp32 partial sum of the fused statistics

template <typename T>
__device__ __forceinline__ T pr_pick4(const T (&a)[4], int i) {
  return i == 0 ? a[0] : (i == 1 ? a[1] : (i == 2 ? a[2] : a[3]));
}
__device__ __forceinline__ void pr_cp16(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(
                   (unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

template <int CO, int Q, bool MOM>
__global__ void __launch_bounds__(PR_T)
    pw_fwd_ring_kernel(const __grid_constant__ PwRingArgs A, int tiles_per_cta) {
  extern __shared__ __align__(16) float4 pr_ring[];          // [Q][PR_S][PR_T], weights, output patches
  float* Wsm = reinterpret_cast<float*>(pr_ring + Q * PR_S * PR_T);   // [Cin = 4Q][CO]
  // per-warp output patch [32 voxels][CO] (CO > 4, dense destination): see the header
  float4* const opatch = reinterpret_cast<float4*>(Wsm + 4 * Q * CO) + (threadIdx.x >> 5) * (32 * CO / 4);
  const bool via_patch = CO > 4 && A.ld_dst == CO;
  __shared__ double sm_mom[2 * CO];
  constexpr int CIN = 4 * Q;
  if (threadIdx.x < 2 * CO) sm_mom[threadIdx.x] = 0.0;
  for (int i = threadIdx.x; i < CIN * CO; i += PR_T) {
    const int j = i % CO, ci = i / CO;
    Wsm[i] = j < A.Cout ? __ldg(A.w + (long long)j * CIN + ci) : 0.f;
  }
  __syncthreads();

  const unsigned ntiles = (A.nvox + PR_T - 1) / PR_T;
  const unsigned tile0 = blockIdx.x * (unsigned)tiles_per_cta;
  const unsigned tile1 = min(tile0 + (unsigned)tiles_per_cta, ntiles);
  const unsigned vend = min(tile1 * PR_T, A.nvox);
  unsigned v = tile0 * PR_T + threadIdx.x;       // consume side
  unsigned vi = v;                               // issue side

  // running source pointers, one per channel group
  const float* px[Q];
  long long sx[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int c = 4 * q;
    const int sg = c / A.seg_w, off = c - sg * A.seg_w;
    const int ld = pr_pick4(A.ld_src, sg);
    px[q] = pr_pick4(A.src, sg) + off + (long long)vi * ld;
    sx[q] = (long long)PR_T * ld;
  }
  float* pdst = A.dst + (long long)v * A.ld_dst;
  const long long sdst = (long long)PR_T * A.ld_dst;

  float4* const r0 = pr_ring + threadIdx.x;
  float4* const r_end = r0 + PR_S * PR_T;
  float4* is = r0;
  auto issue = [&]() {
    if (vi < vend) {
#pragma unroll
      for (int q = 0; q < Q; ++q) pr_cp16(is + q * PR_S * PR_T, px[q]);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    vi += PR_T;
#pragma unroll
    for (int q = 0; q < Q; ++q) px[q] += sx[q];
    is += PR_T;
    if (is == r_end) is = r0;
  };
#pragma unroll
  for (int s = 0; s < PR_S - 1; ++s) issue();

  // per-(n, cin) prologue scale of the current sample (uniform over the CTA: Vs % PR_T == 0)
  const bool has_scale = A.scale != nullptr;
  float4 sc[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) sc[q] = make_float4(1.f, 1.f, 1.f, 1.f);
  int cur_n = -1;
  float ms[MOM ? CO : 1], mq[MOM ? CO : 1];
#pragma unroll
  for (int j = 0; j < (MOM ? CO : 1); ++j) ms[j] = mq[j] = 0.f;
  int since_flush = 0;

  // fp32 partials -> warp sum -> fp64 per-CTA array
  auto partials_to_cta = [&]() {
#pragma unroll
    for (int j = 0; j < (MOM ? CO : 1); ++j) {
      const float s = warp_sum(ms[j]);
      const float q2 = warp_sum(mq[j]);
      if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sm_mom[2 * j + 0], (double)s);
        atomicAdd(&sm_mom[2 * j + 1], (double)q2);
      }
      ms[j] = mq[j] = 0.f;
    }
    since_flush = 0;
  };
  // per-CTA array -> global moments of sample n (all threads of the CTA call this together)
  auto cta_to_global = [&](int n) {
    __syncthreads();
    if (threadIdx.x < 2 * CO && (int)(threadIdx.x >> 1) < A.Cout) {
      atomicAdd(&A.moments[((long long)n * A.Cout) * 2 + threadIdx.x], sm_mom[threadIdx.x]);
      sm_mom[threadIdx.x] = 0.0;
    }
    __syncthreads();
  };

  const float4* cs = r0;
#pragma unroll 1
  for (unsigned t = tile0; t < tile1; ++t) {
    issue();
    asm volatile("cp.async.wait_group %0;\n" ::"n"(PR_S - 1) : "memory");
    // sample of this tile (uniform over the CTA)
    const int tn = (int)((t * PR_T) / A.Vs);
    if (tn != cur_n) {
      if (MOM && cur_n >= 0) {
        partials_to_cta();
        cta_to_global(cur_n);
      }
      cur_n = tn;
      if (has_scale) {
#pragma unroll
        for (int q = 0; q < Q; ++q) sc[q] = ldg4(A.scale + (long long)tn * CIN + 4 * q);
      }
    }
    if (v < vend) {
      float2 acc[CO / 2];
#pragma unroll
      for (int j = 0; j < CO / 2; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        float4 x4 = cs[q * PR_S * PR_T];
        if (A.relu) {
          x4.x = fmaxf(x4.x, 0.f); x4.y = fmaxf(x4.y, 0.f);
          x4.z = fmaxf(x4.z, 0.f); x4.w = fmaxf(x4.w, 0.f);
        }
        if (has_scale) { x4.x *= sc[q].x; x4.y *= sc[q].y; x4.z *= sc[q].z; x4.w *= sc[q].w; }
        const float xe[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float* wr = Wsm + (4 * q + e) * CO;
#pragma unroll
          for (int j4 = 0; j4 < CO / 4; ++j4) {
            const float4 w4 = *reinterpret_cast<const float4*>(wr + 4 * j4);
            fma2(acc[2 * j4 + 0], xe[e], w4.x, w4.y);
            fma2(acc[2 * j4 + 1], xe[e], w4.z, w4.w);
          }
        }
      }
#pragma unroll
      for (int j4 = 0; j4 < CO / 4; ++j4) {
        float o[4] = {acc[2 * j4].x, acc[2 * j4].y, acc[2 * j4 + 1].x, acc[2 * j4 + 1].y};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int cj = 4 * j4 + e;
          if (A.bias && cj < A.Cout) o[e] += __ldg(A.bias + cj);
          if (A.sigmoid) o[e] = 1.f / (1.f + __expf(-o[e]));
          if (MOM) { ms[cj] += o[e]; mq[cj] += o[e] * o[e]; }
        }
        if (via_patch) opatch[(threadIdx.x & 31) * (CO / 4) + j4] = make_float4(o[0], o[1], o[2], o[3]);
        else if (4 * j4 < A.ld_dst) st4(pdst + 4 * j4, make_float4(o[0], o[1], o[2], o[3]));
      }
    }
    if (via_patch) {       // warp-uniform; tiles are whole warps of voxels except at the very end
      __syncwarp();
      const unsigned lane = threadIdx.x & 31;
      const unsigned wv0 = v - lane;                               // first voxel of this warp's patch
      float* wdst = pdst - (long long)lane * CO;
#pragma unroll
      for (int j = 0; j < CO / 4; ++j) {
        const unsigned idx = lane + 32 * j;                        // float4 index inside the patch
        if (wv0 + idx / (CO / 4) < vend) st4(wdst + 4 * idx, opatch[idx]);
      }
      __syncwarp();
    }
    if (MOM && ++since_flush == PR_FLUSH) partials_to_cta();
    v += PR_T;
    pdst += sdst;
    cs += PR_T;
    if (cs == r_end) cs = r0;
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  if (MOM && cur_n >= 0) {
    partials_to_cta();
    cta_to_global(cur_n);
  }
}

template <int CO, int Q, bool MOM>
static int pr_launch(const PwRingArgs& A, cudaStream_t st) {
  auto kern = pw_fwd_ring_kernel<CO, Q, MOM>;
  const int smem = Q * PR_S * PR_T * (int)sizeof(float4) + 4 * Q * CO * (int)sizeof(float) +
                   (CO > 4 ? PR_T * CO * (int)sizeof(float) : 0);
  static int resident = -1;
  if (resident < 0) {
    NAS3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int o = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, PR_T, smem) != cudaSuccess || o < 1) o = 2;
    resident = o * kNumSMs;
  }
  const unsigned ntiles = (A.nvox + PR_T - 1) / PR_T;
  unsigned gx = (unsigned)resident;
  if (gx > ntiles) gx = ntiles;
  if (gx < 1) gx = 1;
  const unsigned tpc = (ntiles + gx - 1) / gx;
  gx = (ntiles + tpc - 1) / tpc;
  kern<<<gx, PR_T, smem, st>>>(A, (int)tpc);
  return NAS3D_OK;
}

template <int CO, int Q>
static int pr_launch_m(const PwRingArgs& A, cudaStream_t st) {
  return A.moments ? pr_launch<CO, Q, true>(A, st) : pr_launch<CO, Q, false>(A, st);
}

// NAS3D_ERR_UNSUPPORTED (no error text) = "not my shape / not enabled": the caller uses pointwise_kernel
int pointwise_sfb_ring(const nas3d_conv_desc* d, const float* big, const float* w, const float* bias,
                       const float* scale, int relu, int sigmoid, float* small, double* moments,
                       cudaStream_t st, const PwCat* cat) {
  // measured on B200: conv1x1_cat_fwd 0.79 -> 0.70 ms per searched-net step
  // (profiles/r2a_ab_optin_variants.txt); option pw_fwd_ring = 0: pointwise_kernel
  if (!g_opt.pw_fwd_ring) return NAS3D_ERR_UNSUPPORTED;
  if (d->k != 1 || d->stride != 1 || d->pad != 0 || d->depthwise) return NAS3D_ERR_UNSUPPORTED;
  if (d->Db != d->Ds || d->Hb != d->Hs || d->Wb != d->Ws) return NAS3D_ERR_UNSUPPORTED;
  const long long nvox = (long long)d->N * d->Ds * d->Hs * d->Ws;
  const long long Vs = (long long)d->Ds * d->Hs * d->Ws;
  if (nvox < (1ll << 20) || nvox >= (1ll << 31) - PR_T) return NAS3D_ERR_UNSUPPORTED;   // big tensors only
  if (Vs % PR_T != 0) return NAS3D_ERR_UNSUPPORTED;       // tiles must not straddle samples
  if (d->Cb % 4 || d->ld_small % 4 || !aligned16(small)) return NAS3D_ERR_UNSUPPORTED;
  const int co_pad = (d->Cs + 3) / 4 * 4;
  if (co_pad > d->ld_small) return NAS3D_ERR_UNSUPPORTED; // whole float4s are stored
  PwRingArgs A{};
  const int nparts = cat ? cat->nparts : 1;
  if (nparts < 1 || nparts > 4 || d->Cb % nparts || (d->Cb / nparts) % 4) return NAS3D_ERR_UNSUPPORTED;
  A.nseg = nparts; A.seg_w = d->Cb / nparts;
  for (int i = 0; i < nparts; ++i) {
    A.src[i] = cat ? cat->src[i] : big;
    A.ld_src[i] = cat ? cat->ld[i] : d->ld_big;
    if (!A.src[i] || !aligned16(A.src[i]) || A.ld_src[i] % 4) return NAS3D_ERR_UNSUPPORTED;
  }
  if (scale && !aligned16(scale)) return NAS3D_ERR_UNSUPPORTED;
  A.w = w; A.bias = bias; A.scale = scale; A.dst = small; A.moments = moments;
  A.Cin = d->Cb; A.Cout = d->Cs; A.ld_dst = d->ld_small;
  A.nvox = (unsigned)nvox; A.Vs = (unsigned)Vs;
  A.relu = relu; A.sigmoid = sigmoid;
  const int Q = d->Cb / 4;
  int rc = NAS3D_ERR_UNSUPPORTED;
  if (co_pad == 4) {
    if (Q == 1) rc = pr_launch_m<4, 1>(A, st);
    else if (Q == 3) rc = pr_launch_m<4, 3>(A, st);
    else if (Q == 6) rc = pr_launch_m<4, 6>(A, st);
  } else if (co_pad == 8) {
    if (Q == 3) rc = pr_launch_m<8, 3>(A, st);
    else if (Q == 6) rc = pr_launch_m<8, 6>(A, st);
  } else if (co_pad == 12) {
    if (Q == 1) rc = pr_launch_m<12, 1>(A, st);
  }
  if (rc != NAS3D_OK) return rc;
  return launched("pointwise_fwd_ring");
}

}  // namespace nas3d
