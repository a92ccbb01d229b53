// 1x1x1 convolutions (stems, cell preprocess, depthwise-separable pointwise, the 12->3 head):
// per-voxel [Cin]x[Cin,Cout] products on NDHWC fp32.  AI is 1.5-6 FLOP/B, so
// these are HBM-bound streaming kernels: one thread per voxel, the voxel's channels are read with
// 128-bit loads (a warp covers 32 consecutive voxels = one contiguous span), all <=16 outputs of
// a channel tile stay in registers, weights are broadcast from shared memory.
//
// "conv view" as in nas3d_b200.h with k = 1, pad = 0:  big_pos = small_pos * stride.
//   SFB  small[o,cs]  = bias + sum_cb f(big[o*s,cb]) W[cs][cb]      (Conv3d forward)
//   BFS  big[o*s,cb] (+)= sum_cs small[o,cs] W[cs][cb]              (Conv3d dgrad; stride 2 only
//                                                                    touches the even voxels)
//   WG   dW[cs][cb]  += sum_o small[o,cs] f(big[o*s,cb])
#include <stdlib.h>
#include "common.cuh"
#include "conv_tiled.h"

namespace nas3d {

struct PwArgs {
  const float* src;
  const float* w;
  const float* bias;
  const float* scale;   // SFB: prologue per-(n,cin) scale ; BFS: epilogue per-(n,cout) scale
  const float* mask;    // BFS: relu mask source (big tensor), pitch ld_mask
  float* dst;
  int N, Ds, Hs, Ws;    // small extents (threads iterate the small lattice)
  int Db, Hb, Wb;       // big extents
  int Cin, Cout;        // channels of src / dst
  int ld_src, ld_dst, ld_mask;
  int stride;
  int relu, sigmoid, accumulate;
  int w_stride_in, w_stride_out;   // W element (in,out) at w[in*w_stride_in + out*w_stride_out]
  double* moments;                 // optional fused GN statistics of dst (SFB, V % 128 == 0)
  // virtual concat (cell.py:82 without the copy): when nseg > 1 the BIG tensor is nseg parts of
  // seg_w channels each, living in separate dense buffers
  int nseg, seg_w;
  const float* seg_src[4];         // SFB / wgrad: parts of the big tensor
  float* seg_dst[4];               // BFS: parts of the big gradient
  const float* seg_mask[4];        // BFS: parts of the relu-mask source
  int seg_ld[4], seg_mask_ld[4], seg_acc[4];
};

// static-index selection: dynamic indexing of a kernel-parameter array would force a per-thread
// local-memory copy of the whole argument block
template <typename T>
__device__ __forceinline__ T pick4(const T (&a)[4], int i) {
  return i == 0 ? a[0] : (i == 1 ? a[1] : (i == 2 ? a[2] : a[3]));
}

constexpr int PW_T = 128;
constexpr int PW_MAX_W = 4096;   // floats of weights per (Cin x CO tile) in smem

// Persistent forward / dgrad kernel.  The grid is exactly the resident capacity (occupancy x SMs) and
// strides over tiles of PW_T*VPT voxels, so there is no partial last wave (the
// thread's voxels are tile + v*PW_T + tid, so every warp access is one contiguous span); the weight
// tile is staged in shared memory once per CTA; the VPT independent 128-bit loads of a channel group
// are issued back to back before the FMAs that consume them.  Optional fused GroupNorm statistics:
// per tile the thread sums its VPT voxels, the warp reduces in fp32 (<=128 values per partial sum),
// partial sums go to a per-CTA fp64 array that is flushed with one atomic per (channel, moment)
// when the CTA crosses a sample boundary or ends (host guarantees tiles do not straddle samples).
template <int CO, bool SRC_IS_BIG, int VPT, bool MOM>
#ifndef NAS3D_PW_MINB
#define NAS3D_PW_MINB 1
#endif
__global__ void __launch_bounds__(PW_T, NAS3D_PW_MINB) pointwise_kernel(const __grid_constant__ PwArgs A) {
  __shared__ __align__(16) float Wsm[PW_MAX_W];
  __shared__ double sm_mom[2 * CO];
  if (threadIdx.x < 2 * CO) sm_mom[threadIdx.x] = 0.0;
  const int co0 = blockIdx.y * CO;
  // Wsm[ci][CO]
  for (int i = threadIdx.x; i < A.Cin * CO; i += PW_T) {
    const int j = i % CO, ci = i / CO;
    Wsm[i] = (co0 + j < A.Cout) ? __ldg(A.w + (long long)ci * A.w_stride_in + (long long)(co0 + j) * A.w_stride_out) : 0.f;
  }
  __syncthreads();
  constexpr unsigned TILE = PW_T * VPT;
  const unsigned Vs = (unsigned)(A.Ds * A.Hs * A.Ws);
  const unsigned nvox = (unsigned)A.N * Vs;
  const unsigned ntiles = (nvox + TILE - 1) / TILE;
  const bool vec_in = (A.Cin % 4 == 0) && (A.ld_src % 4 == 0);
  const bool vec_out = (A.ld_dst % 4 == 0) && (co0 + CO <= A.Cout);
  // fast epilogue: whole float4 channel groups, all per-channel operands fetched as float4
  // a ragged Cout (the 12->3 head) stored at a multiple-of-4 pitch is written as whole float4s
  // too: the pad lanes receive act(0) and are never read as data
  const bool pad_out = SRC_IS_BIG && (A.ld_dst % 4 == 0) && (co0 + CO <= A.ld_dst);
  const bool fast = ((A.Cout % 4 == 0 && (A.nseg > 1 || vec_out)) || pad_out) &&
                    (A.mask == nullptr || A.ld_mask % 4 == 0);
  const bool need_n = A.stride != 1 || A.scale != nullptr;
  int mom_n = -1;
  float ms[MOM ? CO : 1], mq[MOM ? CO : 1];   // per-thread partial statistics, carried across tiles
#pragma unroll
  for (int j = 0; j < (MOM ? CO : 1); ++j) ms[j] = mq[j] = 0.f;
  // warp-reduce the partials into the CTA's fp64 array, then one atomic per (channel, moment)
  auto flush_moments = [&](int n_flush) {
#pragma unroll
    for (int j = 0; j < (MOM ? CO : 1); ++j) {
      const float a = warp_sum(ms[j]);
      const float b = warp_sum(mq[j]);
      if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sm_mom[j * 2 + 0], (double)a);
        atomicAdd(&sm_mom[j * 2 + 1], (double)b);
      }
      ms[j] = mq[j] = 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * CO; i += PW_T)
      if (co0 + i / 2 < A.Cout) {
        atomicAdd(&A.moments[((long long)n_flush * A.Cout + co0) * 2 + i], sm_mom[i]);
        sm_mom[i] = 0.0;
      }
    __syncthreads();
  };

  for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    unsigned o[VPT];
    long long sidx[VPT], didx[VPT];
    int n[VPT];
    bool ok[VPT];
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      o[v] = tile * TILE + v * PW_T + threadIdx.x;
      ok[v] = o[v] < nvox;
      const unsigned oo = ok[v] ? o[v] : 0u;
      long long big_idx = oo;
      n[v] = 0;
      if (need_n) {
        const unsigned nn = oo / Vs;
        n[v] = (int)nn;
        if (A.stride != 1) {
          unsigned r = oo - nn * Vs;
          const unsigned ow = r % (unsigned)A.Ws; r /= (unsigned)A.Ws;
          const unsigned oh = r % (unsigned)A.Hs;
          const unsigned od = r / (unsigned)A.Hs;
          big_idx = (((long long)nn * A.Db + od * A.stride) * A.Hb + oh * A.stride) * A.Wb + ow * A.stride;
        }
      }
      sidx[v] = SRC_IS_BIG ? big_idx : (long long)oo;
      didx[v] = SRC_IS_BIG ? (long long)oo : big_idx;
    }

    float2 acc[VPT][CO / 2];      // channel pairs: FFMA2 accumulators (common.cuh)
#pragma unroll
    for (int v = 0; v < VPT; ++v)
#pragma unroll
      for (int j = 0; j < CO / 2; ++j) acc[v][j] = make_float2(0.f, 0.f);

#pragma unroll 2
    for (int c4 = 0; c4 < A.Cin; c4 += 4) {
      const int nv = min(4, A.Cin - c4);
      float xv[VPT][4];
      if (SRC_IS_BIG && A.nseg > 1) {
        const int sg = (c4 >= A.seg_w) + (c4 >= 2 * A.seg_w) + (c4 >= 3 * A.seg_w);
        const float* sb = pick4(A.seg_src, sg) + (c4 - sg * A.seg_w);
        const int sl = pick4(A.seg_ld, sg);
#pragma unroll
        for (int v = 0; v < VPT; ++v) {
          const float4 t = ldg4(sb + sidx[v] * sl);
          xv[v][0] = t.x; xv[v][1] = t.y; xv[v][2] = t.z; xv[v][3] = t.w;
        }
      } else if (vec_in) {
#pragma unroll
        for (int v = 0; v < VPT; ++v) {
          const float4 t = ldg4(A.src + sidx[v] * A.ld_src + c4);
          xv[v][0] = t.x; xv[v][1] = t.y; xv[v][2] = t.z; xv[v][3] = t.w;
        }
      } else {
#pragma unroll
        for (int v = 0; v < VPT; ++v)
#pragma unroll
          for (int e = 0; e < 4; ++e)
            xv[v][e] = e < nv ? __ldg(A.src + sidx[v] * A.ld_src + c4 + e) : 0.f;
      }
      if (SRC_IS_BIG) {
        if (A.relu) {
#pragma unroll
          for (int v = 0; v < VPT; ++v)
#pragma unroll
            for (int e = 0; e < 4; ++e) xv[v][e] = fmaxf(xv[v][e], 0.f);
        }
        if (A.scale) {
#pragma unroll
          for (int v = 0; v < VPT; ++v)
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (e < nv) xv[v][e] *= __ldg(A.scale + (long long)n[v] * A.Cin + c4 + e);
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (e < nv) {
          const float* wr = Wsm + (c4 + e) * CO;
#pragma unroll
          for (int j4 = 0; j4 < CO / 4; ++j4) {
            const float4 w4 = *reinterpret_cast<const float4*>(wr + j4 * 4);
#pragma unroll
            for (int v = 0; v < VPT; ++v) {
              fma2(acc[v][j4 * 2 + 0], xv[v][e], w4.x, w4.y);
              fma2(acc[v][j4 * 2 + 1], xv[v][e], w4.z, w4.w);
            }
          }
        }
      }
    }

    if (MOM) {   // uniform per CTA: all voxels of a tile belong to one sample
      const int tn = (int)((tile * TILE) / Vs);
      if (tn != mom_n) {
        if (mom_n >= 0) flush_moments(mom_n);
        mom_n = tn;
      }
    }

#pragma unroll
    for (int j4 = 0; j4 < CO / 4; ++j4) {
      const int c = co0 + j4 * 4;
      if (fast && c >= A.Cout) continue;
      // per-channel operands shared by the thread's voxels
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      int sg = 0, off = c;
      if (fast) {
        // segment of this channel group (equal widths, multiples of 4): no integer division
        sg = (A.nseg > 1) ? (c >= A.seg_w) + (c >= 2 * A.seg_w) + (c >= 3 * A.seg_w) : 0;
        off = c - sg * A.seg_w;
        if (A.bias) {
          if (c + 4 <= A.Cout) {
            b4 = ldg4(A.bias + c);
          } else {
            b4.x = __ldg(A.bias + c);
            if (c + 1 < A.Cout) b4.y = __ldg(A.bias + c + 1);
            if (c + 2 < A.Cout) b4.z = __ldg(A.bias + c + 2);
          }
        }
      }
      // relu-mask operands of the thread's voxels: issued together, consumed below
      float4 m4[VPT];
      bool has_mask = false;
      if (fast && !SRC_IS_BIG) {
        const float* mb = A.nseg > 1 ? pick4(A.seg_mask, sg) : A.mask;
        const int ml = A.nseg > 1 ? pick4(A.seg_mask_ld, sg) : A.ld_mask;
        const int mo = A.nseg > 1 ? off : c;
        has_mask = mb != nullptr;
        if (has_mask) {
#pragma unroll
          for (int vv = 0; vv < VPT; ++vv) m4[vv] = ldg4(mb + didx[vv] * ml + mo);
        }
      }
#pragma unroll
      for (int vv = 0; vv < VPT; ++vv) {
        if (!ok[vv]) continue;
        float v[4] = {acc[vv][j4 * 2].x, acc[vv][j4 * 2].y, acc[vv][j4 * 2 + 1].x, acc[vv][j4 * 2 + 1].y};
        const long long dst_idx = didx[vv];
        float* pd = A.dst + dst_idx * A.ld_dst + co0;
        if (fast) {
          v[0] += b4.x; v[1] += b4.y; v[2] += b4.z; v[3] += b4.w;
          if (!SRC_IS_BIG) {
            if (has_mask) {
              v[0] = m4[vv].x > 0.f ? v[0] : 0.f; v[1] = m4[vv].y > 0.f ? v[1] : 0.f;
              v[2] = m4[vv].z > 0.f ? v[2] : 0.f; v[3] = m4[vv].w > 0.f ? v[3] : 0.f;
            }
            if (A.scale) {
              const float4 s4 = ldg4(A.scale + (long long)n[vv] * A.Cout + c);
              v[0] *= s4.x; v[1] *= s4.y; v[2] *= s4.z; v[3] *= s4.w;
            }
          }
          if (A.sigmoid) {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = 1.f / (1.f + __expf(-v[e]));
          }
          if (MOM) {
#pragma unroll
            for (int e = 0; e < 4; ++e) { ms[j4 * 4 + e] += v[e]; mq[j4 * 4 + e] += v[e] * v[e]; }
          }
          float* ps;
          int do_acc;
          if (!SRC_IS_BIG && A.nseg > 1) {
            ps = pick4(A.seg_dst, sg) + dst_idx * pick4(A.seg_ld, sg) + off;
            do_acc = pick4(A.seg_acc, sg);
          } else {
            ps = pd + j4 * 4;
            do_acc = A.accumulate;
          }
          float4 r = make_float4(v[0], v[1], v[2], v[3]);
          if (do_acc) {
            const float4 old = *reinterpret_cast<const float4*>(ps);
            r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w;
          }
          st4(ps, r);
          continue;
        }
        // generic scalar path (Cout not a multiple of 4, e.g. the 12->3 head)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int ce = c + e;
          float t = v[e];
          if (ce < A.Cout) {
            if (A.bias) t += __ldg(A.bias + ce);
            if (!SRC_IS_BIG) {
              if (A.mask) t = (__ldg(A.mask + dst_idx * A.ld_mask + ce) > 0.f) ? t : 0.f;
              if (A.scale) t *= __ldg(A.scale + (long long)n[vv] * A.Cout + ce);
            }
            if (A.sigmoid) t = 1.f / (1.f + __expf(-t));
            if (MOM) { ms[j4 * 4 + e] += t; mq[j4 * 4 + e] += t * t; }
            pd[j4 * 4 + e] = A.accumulate ? pd[j4 * 4 + e] + t : t;
          }
        }
      }
    }
  }
  if (MOM && mom_n >= 0) flush_moments(mom_n);
}

// wgrad: thread register tile TS x TB over (small channels, big channels)
constexpr int PWG_T = 256;
constexpr int PWG_ITER = 128;   // voxels per thread before the 5-round warp reduction of its tile

template <int TS, int TB>
__global__ void __launch_bounds__(PWG_T)
    pointwise_wgrad_kernel(const __grid_constant__ PwArgs A, const float* __restrict__ small,
                           const float* __restrict__ big, int Cs, int Cb, int lds, int ldb,
                           float* __restrict__ dW, float* __restrict__ dbias_small, int iters) {
  const int cs0 = blockIdx.y * TS, cb0 = blockIdx.z * TB;
  const int ns = min(TS, Cs - cs0), nb = min(TB, Cb - cb0);
  const bool vec_s = (lds % 4 == 0) && (ns == TS) && (TS % 4 == 0);
  const bool vec_b = (ldb % 4 == 0) && (nb == TB);
  const long long nvox = (long long)A.N * A.Ds * A.Hs * A.Ws;
  static_assert(TB % 2 == 0, "big-channel tile is processed in FFMA2 pairs");
  float2 acc[TS][TB / 2];
  float bs[TS];
#pragma unroll
  for (int i = 0; i < TS; ++i) {
    bs[i] = 0.f;
#pragma unroll
    for (int j = 0; j < TB / 2; ++j) acc[i][j] = make_float2(0.f, 0.f);
  }
  const bool do_bias = dbias_small != nullptr && blockIdx.z == 0;
  const bool need_decomp = A.stride != 1 || A.scale != nullptr;    // host guarantees nvox < 2^31
  // trip count known up front (no early exit in the body), so the loads of the next voxels can
  // be hoisted above the FMAs of the current one
  const long long o0 = (long long)blockIdx.x * iters * PWG_T + threadIdx.x;
  const int my_iters = o0 < nvox ? (int)min((long long)iters, (nvox - o0 + PWG_T - 1) / PWG_T) : 0;
#pragma unroll 4
  for (int it = 0; it < my_iters; ++it) {
    const long long o = o0 + (long long)it * PWG_T;
    long long bidx = o;
    int n = 0;
    if (need_decomp) {
      // 32-bit voxel decomposition, only where the sample / big-lattice index is needed at all
      // (the 64-bit div/mod chain cost more instructions than the FMAs of the voxel)
      unsigned t = (unsigned)o;
      const unsigned ow = t % (unsigned)A.Ws; t /= (unsigned)A.Ws;
      const unsigned oh = t % (unsigned)A.Hs; t /= (unsigned)A.Hs;
      const unsigned od = t % (unsigned)A.Ds;
      n = (int)(t / (unsigned)A.Ds);
      if (A.stride != 1)
        bidx = (((long long)n * A.Db + od * A.stride) * A.Hb + oh * A.stride) * A.Wb + ow * A.stride;
    }
    float sv[TS], bv[TB];
    const float* ps = small + o * lds + cs0;
    const float* pb = big + bidx * ldb + cb0;
    if (vec_s) {
#pragma unroll
      for (int i = 0; i < TS; i += 4) {
        const float4 t = ldg4(ps + i);
        sv[i] = t.x; sv[i + 1] = t.y; sv[i + 2] = t.z; sv[i + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < TS; ++i) sv[i] = i < ns ? __ldg(ps + i) : 0.f;
    }
    if (A.nseg > 1) {
#pragma unroll
      for (int j = 0; j < TB; j += 4) {
        const int c = cb0 + j;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < Cb) {
          const int sg = (c >= A.seg_w) + (c >= 2 * A.seg_w) + (c >= 3 * A.seg_w);
          t = ldg4(pick4(A.seg_src, sg) + bidx * pick4(A.seg_ld, sg) + (c - sg * A.seg_w));
        }
        bv[j] = t.x; bv[j + 1] = t.y; bv[j + 2] = t.z; bv[j + 3] = t.w;
      }
    } else if (vec_b) {
#pragma unroll
      for (int j = 0; j < TB; j += 4) {
        const float4 t = ldg4(pb + j);
        bv[j] = t.x; bv[j + 1] = t.y; bv[j + 2] = t.z; bv[j + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < TB; ++j) bv[j] = j < nb ? __ldg(pb + j) : 0.f;
    }
    if (A.relu) {
#pragma unroll
      for (int j = 0; j < TB; ++j) bv[j] = fmaxf(bv[j], 0.f);
    }
    if (A.scale) {
#pragma unroll
      for (int j = 0; j < TB; ++j)
        if (j < nb) bv[j] *= __ldg(A.scale + (long long)n * Cb + cb0 + j);
    }
#pragma unroll
    for (int i = 0; i < TS; ++i) {
      bs[i] += sv[i];
#pragma unroll
      for (int j = 0; j < TB / 2; ++j) fma2(acc[i][j], sv[i], bv[2 * j], bv[2 * j + 1]);
    }
  }
  // warp totals of the TS*TB products and TS bias sums with one transposing butterfly
  constexpr int NV = TS * TB + TS, NVP = (NV + 31) / 32 * 32, PER = NVP / 32;
  __shared__ float red[PWG_T / 32][NVP];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  {
    float vals[NVP];
#pragma unroll
    for (int i = 0; i < TS; ++i) {
#pragma unroll
      for (int j = 0; j < TB; ++j) vals[i * TB + j] = (j & 1) ? acc[i][j >> 1].y : acc[i][j >> 1].x;
      vals[TS * TB + i] = bs[i];
    }
#pragma unroll
    for (int i = NV; i < NVP; ++i) vals[i] = 0.f;
    warp_reduce_scatter<NVP>(vals);
#pragma unroll
    for (int j = 0; j < PER; ++j) red[wid][lane * PER + j] = vals[j];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < TS * TB + TS; e += PWG_T) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < PWG_T / 32; ++w) v += red[w][e];
    if (e < TS * TB) {
      const int i = e / TB, j = e % TB;
      if (i < ns && j < nb) atomicAdd(dW + (long long)(cs0 + i) * Cb + cb0 + j, v);
    } else if (do_bias) {
      const int i = e - TS * TB;
      if (i < ns) atomicAdd(dbias_small + cs0 + i, v);
    }
  }
}

static inline bool pointwise_big(long long nvox) { return nvox >= (long long)kNumSMs * 8 * PW_T * 4; }

template <int CO, bool SRC_IS_BIG, int VPT, bool MOM>
static unsigned pointwise_resident_ctas() {
  static int occ = 0;   // per instantiation; racing initialisers compute the same value
  if (occ == 0) {
    int o = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, pointwise_kernel<CO, SRC_IS_BIG, VPT, MOM>, PW_T, 0) != cudaSuccess || o < 1)
      o = 4;
    occ = o;
  }
  return (unsigned)occ * kNumSMs;
}

template <int CO, bool SRC_IS_BIG, bool MOM, int VPT>
static void launch_pointwise_vpt(const PwArgs& A, unsigned gy, long long nvox, cudaStream_t st) {
  const unsigned ntiles = (unsigned)((nvox + PW_T * VPT - 1) / (PW_T * VPT));
  unsigned gx = pointwise_resident_ctas<CO, SRC_IS_BIG, VPT, MOM>() / gy;
  if (gx < 1) gx = 1;
  if (gx > ntiles) gx = ntiles;
  pointwise_kernel<CO, SRC_IS_BIG, VPT, MOM><<<dim3(gx, gy), PW_T, 0, st>>>(A);
}

template <int CO, bool SRC_IS_BIG, bool MOM>
static void launch_pointwise_co(const PwArgs& A, unsigned gy, cudaStream_t st) {
  const long long nvox = (long long)A.N * A.Ds * A.Hs * A.Ws;
  // big tensors: several voxels per thread, one resident wave of CTAs striding over the tiles;
  // small (deep-level) tensors: 1 voxel per thread, one tile per CTA - they are latency-bound
  if (pointwise_big(nvox)) {
    const int vpt = MOM ? g_opt.pw_vpt_mom : (SRC_IS_BIG ? g_opt.pw_vpt_sfb : g_opt.pw_vpt_bfs);
    if (vpt == 1) launch_pointwise_vpt<CO, SRC_IS_BIG, MOM, 1>(A, gy, nvox, st);
    else if (vpt == 2) launch_pointwise_vpt<CO, SRC_IS_BIG, MOM, 2>(A, gy, nvox, st);
    else launch_pointwise_vpt<CO, SRC_IS_BIG, MOM, 4>(A, gy, nvox, st);
  } else {
    const unsigned ntiles = (unsigned)((nvox + PW_T - 1) / PW_T);
    pointwise_kernel<CO, SRC_IS_BIG, 1, MOM><<<dim3(ntiles, gy), PW_T, 0, st>>>(A);
  }
}

template <bool SRC_IS_BIG>
static int launch_pointwise(const PwArgs& A, cudaStream_t st) {
  const long long nvox = (long long)A.N * A.Ds * A.Hs * A.Ws;
  if (nvox >= (1ll << 31) - PW_T * 4) return NAS3D_ERR_UNSUPPORTED;   // 32-bit voxel indices
  int co_t = A.Cout <= 4 ? 4 : A.Cout <= 8 ? 8 : A.Cout <= 12 ? 12 : 16;
  if (A.Cin * co_t > PW_MAX_W) return NAS3D_ERR_UNSUPPORTED;
  const unsigned gy = (unsigned)((A.Cout + co_t - 1) / co_t);
  if (SRC_IS_BIG && A.moments) {
    switch (co_t) {
      case 4: launch_pointwise_co<4, SRC_IS_BIG, SRC_IS_BIG>(A, gy, st); break;
      case 8: launch_pointwise_co<8, SRC_IS_BIG, SRC_IS_BIG>(A, gy, st); break;
      case 12: launch_pointwise_co<12, SRC_IS_BIG, SRC_IS_BIG>(A, gy, st); break;
      default: launch_pointwise_co<16, SRC_IS_BIG, SRC_IS_BIG>(A, gy, st); break;
    }
    return launched("pointwise");
  }
  switch (co_t) {
    case 4: launch_pointwise_co<4, SRC_IS_BIG, false>(A, gy, st); break;
    case 8: launch_pointwise_co<8, SRC_IS_BIG, false>(A, gy, st); break;
    case 12: launch_pointwise_co<12, SRC_IS_BIG, false>(A, gy, st); break;
    default: launch_pointwise_co<16, SRC_IS_BIG, false>(A, gy, st); break;
  }
  return launched("pointwise");
}

// entry points used by conv_direct.cu; NAS3D_ERR_UNSUPPORTED = "not my shape" (no error text)
static bool fill_segments(PwArgs* A, const PwCat* cat, int Cb) {
  A->nseg = 1; A->seg_w = Cb;
  if (!cat || cat->nparts <= 1) return true;
  if (cat->nparts > 4 || Cb % cat->nparts) return false;
  A->nseg = cat->nparts; A->seg_w = Cb / cat->nparts;
  if (A->seg_w % 4) return false;
  for (int i = 0; i < cat->nparts; ++i) {
    A->seg_src[i] = cat->src ? cat->src[i] : nullptr;
    A->seg_dst[i] = cat->dst ? cat->dst[i] : nullptr;
    A->seg_mask[i] = cat->mask ? cat->mask[i] : nullptr;
    A->seg_ld[i] = cat->ld[i];
    A->seg_mask_ld[i] = cat->mask_ld ? cat->mask_ld[i] : 0;
    A->seg_acc[i] = cat->acc ? cat->acc[i] : 0;
    if (A->seg_ld[i] % 4) return false;
    if (A->seg_src[i] && !aligned16(A->seg_src[i])) return false;
    if (A->seg_dst[i] && !aligned16(A->seg_dst[i])) return false;
  }
  return true;
}

int pointwise_sfb(const nas3d_conv_desc* d, const float* big, const float* w, const float* bias,
                  const float* scale, int relu, int sigmoid, float* small, int accumulate,
                  double* moments, cudaStream_t st, const PwCat* cat) {
  if (!accumulate) {   // opt-in ring-staged kernel for the big stride-1 shapes (NAS3D_PW_FWD_RING=1)
    const int rc = pointwise_sfb_ring(d, big, w, bias, scale, relu, sigmoid, small, moments, st, cat);
    if (rc != NAS3D_ERR_UNSUPPORTED) return rc;
  }
  // fused statistics need tiles that do not straddle samples
  if (moments) {
    const long long Vs = (long long)d->Ds * d->Hs * d->Ws;
    const long long tile = pointwise_big(Vs * d->N) ? PW_T * 4 : PW_T;
    if (Vs % tile != 0 || accumulate || sigmoid) return NAS3D_ERR_UNSUPPORTED;
  }
  if (!cat && d->Cb % 4 == 0 && d->ld_big % 4 == 0 && !aligned16(big)) return NAS3D_ERR_UNSUPPORTED;
  PwArgs A{};
  if (!fill_segments(&A, cat, d->Cb)) return NAS3D_ERR_UNSUPPORTED;
  A.src = big; A.w = w; A.bias = bias; A.scale = scale; A.mask = nullptr; A.dst = small;
  A.N = d->N; A.Ds = d->Ds; A.Hs = d->Hs; A.Ws = d->Ws; A.Db = d->Db; A.Hb = d->Hb; A.Wb = d->Wb;
  A.Cin = d->Cb; A.Cout = d->Cs; A.ld_src = d->ld_big; A.ld_dst = d->ld_small; A.ld_mask = 0;
  A.stride = d->stride; A.relu = relu; A.sigmoid = sigmoid; A.accumulate = accumulate;
  A.w_stride_in = 1; A.w_stride_out = d->Cb;      // W[cs][cb]: in = cb, out = cs
  A.moments = moments;
  if (d->ld_small % 4 == 0 && !aligned16(small)) return NAS3D_ERR_UNSUPPORTED;
  return launch_pointwise<true>(A, st);
}

int pointwise_bfs(const nas3d_conv_desc* d, const float* small, const float* w, const float* bias,
                  const float* mask_big, int ld_mask, const float* scale, float* big,
                  int accumulate, cudaStream_t st, const PwCat* cat) {
  if (bias) return NAS3D_ERR_UNSUPPORTED;   // a 1x1 transposed conv forward is not in this network
  if (!cat && d->ld_big % 4 == 0 && !aligned16(big)) return NAS3D_ERR_UNSUPPORTED;
  if (d->Cs % 4 == 0 && d->ld_small % 4 == 0 && !aligned16(small)) return NAS3D_ERR_UNSUPPORTED;
  PwArgs A{};
  if (!fill_segments(&A, cat, d->Cb)) return NAS3D_ERR_UNSUPPORTED;
  if (d->stride != 1) {
    // only the stride-lattice voxels receive a value: the rest of a fresh gradient is zero
    const size_t nvb = (size_t)d->N * d->Db * d->Hb * d->Wb;
    if (A.nseg > 1) {
      for (int i = 0; i < A.nseg; ++i)
        if (!A.seg_acc[i]) {
          if (A.seg_ld[i] != A.seg_w) return NAS3D_ERR_UNSUPPORTED;
          NAS3D_CUDA(cudaMemsetAsync(A.seg_dst[i], 0, sizeof(float) * nvb * A.seg_w, st));
        }
    } else if (!accumulate) {
      if (d->ld_big != d->Cb) return NAS3D_ERR_UNSUPPORTED;
      NAS3D_CUDA(cudaMemsetAsync(big, 0, sizeof(float) * nvb * d->Cb, st));
    }
  }
  A.src = small; A.w = w; A.bias = nullptr; A.scale = scale; A.mask = mask_big; A.dst = big;
  A.N = d->N; A.Ds = d->Ds; A.Hs = d->Hs; A.Ws = d->Ws; A.Db = d->Db; A.Hb = d->Hb; A.Wb = d->Wb;
  A.Cin = d->Cs; A.Cout = d->Cb; A.ld_src = d->ld_small; A.ld_dst = d->ld_big; A.ld_mask = ld_mask;
  A.stride = d->stride; A.relu = 0; A.sigmoid = 0; A.accumulate = accumulate;
  A.w_stride_in = d->Cb; A.w_stride_out = 1;      // in = cs, out = cb
  return launch_pointwise<false>(A, st);
}

int pointwise_wgrad(const nas3d_conv_desc* d, const float* small, const float* big,
                    const float* scale, int relu, float* dW, float* dbias_small, cudaStream_t st,
                    const PwCat* cat) {
  if ((!cat && d->ld_big % 4 == 0 && !aligned16(big)) || (d->ld_small % 4 == 0 && !aligned16(small)))
    return NAS3D_ERR_UNSUPPORTED;
  PwArgs A{};
  if (!fill_segments(&A, cat, d->Cb)) return NAS3D_ERR_UNSUPPORTED;
  A.N = d->N; A.Ds = d->Ds; A.Hs = d->Hs; A.Ws = d->Ws; A.Db = d->Db; A.Hb = d->Hb; A.Wb = d->Wb;
  A.stride = d->stride; A.relu = relu; A.scale = scale;
  const long long nvox = (long long)d->N * d->Ds * d->Hs * d->Ws;
  if (nvox >= (1ll << 31)) return NAS3D_ERR_UNSUPPORTED;   // 32-bit voxel indices
  const int Cs = d->Cs, Cb = d->Cb;
  // register tile over (small, big) channels: cover the whole matrix when it is small so each
  // tensor is read once; voxels per thread shrink for small (deep-level) tensors so that at least
  // ~4 CTAs per SM share the reduction
  int ts, tb;
  if (Cs % 12 == 0 && Cb == 4) { ts = 12; tb = 4; }
  else if (Cb % 12 == 0 && Cb <= 48) { ts = 4; tb = 12; }
  else if (Cs % 8 == 0 && Cb % 8 == 0) { ts = 8; tb = 8; }
  else if (Cb % 8 == 0) { ts = 4; tb = 8; }
  else { ts = 4; tb = 4; }
  const unsigned gy = (unsigned)((Cs + ts - 1) / ts), gz = (unsigned)((Cb + tb - 1) / tb);
  long long it = nvox * gy * gz / ((long long)kNumSMs * 4 * PWG_T);
  const int iters = (int)(it < 1 ? 1 : (it > PWG_ITER ? PWG_ITER : it));
  const long long per_block = (long long)PWG_T * iters;
  const unsigned gx = (unsigned)((nvox + per_block - 1) / per_block);
  const dim3 grid(gx, gy, gz);
#define NAS3D_PWG(TS_, TB_)                                                                       \
  pointwise_wgrad_kernel<TS_, TB_><<<grid, PWG_T, 0, st>>>(A, small, big, Cs, Cb, d->ld_small,     \
                                                           d->ld_big, dW, dbias_small, iters)
  if (ts == 12) NAS3D_PWG(12, 4);
  else if (tb == 12) NAS3D_PWG(4, 12);
  else if (ts == 8) NAS3D_PWG(8, 8);
  else if (tb == 8) NAS3D_PWG(4, 8);
  else NAS3D_PWG(4, 4);
#undef NAS3D_PWG
  return launched("pointwise_wgrad");
}

}  // namespace nas3d
