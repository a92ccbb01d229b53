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

// BIGDST: the thread's voxel index runs over the small lattice; src/dst voxel addressing differs
template <int CO, bool SRC_IS_BIG>
__global__ void __launch_bounds__(PW_T) pointwise_kernel(const __grid_constant__ PwArgs A) {
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
  const long long nvox = (long long)A.N * A.Ds * A.Hs * A.Ws;
  const long long o = (long long)blockIdx.x * PW_T + threadIdx.x;
  if (o >= nvox) return;     // never taken when moments are fused (host guarantees V % PW_T == 0)
  long long big_idx = o;
  int n;
  {
    long long t = o;
    const int ow = (int)(t % A.Ws); t /= A.Ws;
    const int oh = (int)(t % A.Hs); t /= A.Hs;
    const int od = (int)(t % A.Ds);
    n = (int)(t / A.Ds);
    if (A.stride != 1)
      big_idx = (((long long)n * A.Db + od * A.stride) * A.Hb + oh * A.stride) * A.Wb + ow * A.stride;
  }
  const long long src_idx = SRC_IS_BIG ? big_idx : o;
  const long long dst_idx = SRC_IS_BIG ? o : big_idx;
  const float* px = A.src + src_idx * A.ld_src;
  const bool vec_in = (A.Cin % 4 == 0) && (A.ld_src % 4 == 0);

  float acc[CO];
#pragma unroll
  for (int j = 0; j < CO; ++j) acc[j] = 0.f;

  for (int c4 = 0; c4 < A.Cin; c4 += 4) {
    const int nv = min(4, A.Cin - c4);
    float xv[4];
    if (SRC_IS_BIG && A.nseg > 1) {
      const int sg = (c4 >= A.seg_w) + (c4 >= 2 * A.seg_w) + (c4 >= 3 * A.seg_w);
      const float4 t = ldg4(pick4(A.seg_src, sg) + src_idx * pick4(A.seg_ld, sg) + (c4 - sg * A.seg_w));
      xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w;
    } else if (vec_in) {
      const float4 t = ldg4(px + c4);
      xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) xv[e] = e < nv ? __ldg(px + c4 + e) : 0.f;
    }
    if (SRC_IS_BIG) {
      if (A.relu) {
#pragma unroll
        for (int e = 0; e < 4; ++e) xv[e] = fmaxf(xv[e], 0.f);
      }
      if (A.scale) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (e < nv) xv[e] *= __ldg(A.scale + (long long)n * A.Cin + c4 + e);
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (e < nv) {
        const float* wr = Wsm + (c4 + e) * CO;
#pragma unroll
        for (int j4 = 0; j4 < CO / 4; ++j4) {
          const float4 w4 = *reinterpret_cast<const float4*>(wr + j4 * 4);
          acc[j4 * 4 + 0] += xv[e] * w4.x; acc[j4 * 4 + 1] += xv[e] * w4.y;
          acc[j4 * 4 + 2] += xv[e] * w4.z; acc[j4 * 4 + 3] += xv[e] * w4.w;
        }
      }
    }
  }

  float* pd = A.dst + dst_idx * A.ld_dst + co0;
  const bool vec_out = (A.ld_dst % 4 == 0) && (co0 + CO <= A.Cout);
  // fast path: whole float4 channel groups, all per-channel operands fetched as float4
  const bool fast = (A.Cout % 4 == 0) && (A.nseg > 1 || vec_out) &&
                    (A.mask == nullptr || A.ld_mask % 4 == 0);
#pragma unroll
  for (int j4 = 0; j4 < CO / 4; ++j4) {
    const int c = co0 + j4 * 4;
    float v[4] = {acc[j4 * 4 + 0], acc[j4 * 4 + 1], acc[j4 * 4 + 2], acc[j4 * 4 + 3]};
    if (fast) {
      if (c >= A.Cout) continue;
      // segment of this channel group (equal widths, multiples of 4): no integer division
      const int sg = (A.nseg > 1) ? (c >= A.seg_w) + (c >= 2 * A.seg_w) + (c >= 3 * A.seg_w) : 0;
      const int off = c - sg * A.seg_w;
      if (A.bias) {
        const float4 b4 = ldg4(A.bias + c);
        v[0] += b4.x; v[1] += b4.y; v[2] += b4.z; v[3] += b4.w;
      }
      if (!SRC_IS_BIG) {
        const float* mp = nullptr;
        if (A.nseg > 1) {
          const float* mb = pick4(A.seg_mask, sg);
          if (mb) mp = mb + dst_idx * pick4(A.seg_mask_ld, sg) + off;
        } else if (A.mask) {
          mp = A.mask + dst_idx * A.ld_mask + c;
        }
        if (mp) {
          const float4 m4 = ldg4(mp);
          v[0] = m4.x > 0.f ? v[0] : 0.f; v[1] = m4.y > 0.f ? v[1] : 0.f;
          v[2] = m4.z > 0.f ? v[2] : 0.f; v[3] = m4.w > 0.f ? v[3] : 0.f;
        }
        if (A.scale) {
          const float4 s4 = ldg4(A.scale + (long long)n * A.Cout + c);
          v[0] *= s4.x; v[1] *= s4.y; v[2] *= s4.z; v[3] *= s4.w;
        }
      }
      if (A.sigmoid) {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = 1.f / (1.f + __expf(-v[e]));
      }
      if (A.moments) {
        float ms[4], mq[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { ms[e] = v[e]; mq[e] = v[e] * v[e]; }
        warp_moments_add(sm_mom, j4 * 4, ms, mq);
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
          if (A.scale) t *= __ldg(A.scale + (long long)n * A.Cout + ce);
        }
        if (A.sigmoid) t = 1.f / (1.f + __expf(-t));
      }
      v[e] = t;
    }
    if (A.moments) {
      float ms[4], mq[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) { ms[e] = v[e]; mq[e] = v[e] * v[e]; }
      warp_moments_add(sm_mom, j4 * 4, ms, mq);
    }
    if (vec_out) {
      float4 r = make_float4(v[0], v[1], v[2], v[3]);
      if (A.accumulate) {
        const float4 old = *reinterpret_cast<const float4*>(pd + j4 * 4);
        r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w;
      }
      st4(pd + j4 * 4, r);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int ce = c + e;
        if (ce < A.Cout) pd[j4 * 4 + e] = A.accumulate ? pd[j4 * 4 + e] + v[e] : v[e];
      }
    }
  }
  if (A.moments) {
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * CO; i += PW_T)
      if (co0 + i / 2 < A.Cout)
        atomicAdd(&A.moments[((long long)n * A.Cout + co0) * 2 + i], sm_mom[i]);
  }
}

// wgrad: thread register tile TS x TB over (small channels, big channels)
constexpr int PWG_T = 256;
constexpr int PWG_ITER = 32;

template <int TS, int TB>
__global__ void __launch_bounds__(PWG_T)
    pointwise_wgrad_kernel(const __grid_constant__ PwArgs A, const float* __restrict__ small,
                           const float* __restrict__ big, int Cs, int Cb, int lds, int ldb,
                           float* __restrict__ dW, float* __restrict__ dbias_small) {
  const int cs0 = blockIdx.y * TS, cb0 = blockIdx.z * TB;
  const int ns = min(TS, Cs - cs0), nb = min(TB, Cb - cb0);
  const bool vec_s = (lds % 4 == 0) && (ns == TS) && (TS % 4 == 0);
  const bool vec_b = (ldb % 4 == 0) && (nb == TB);
  const long long nvox = (long long)A.N * A.Ds * A.Hs * A.Ws;
  float acc[TS][TB];
  float bs[TS];
#pragma unroll
  for (int i = 0; i < TS; ++i) {
    bs[i] = 0.f;
#pragma unroll
    for (int j = 0; j < TB; ++j) acc[i][j] = 0.f;
  }
  const bool do_bias = dbias_small != nullptr && blockIdx.z == 0;
  for (int it = 0; it < PWG_ITER; ++it) {
    const long long o = ((long long)blockIdx.x * PWG_ITER + it) * PWG_T + threadIdx.x;
    if (o >= nvox) break;
    long long bidx = o;
    int n;
    {
      long long t = o;
      const int ow = (int)(t % A.Ws); t /= A.Ws;
      const int oh = (int)(t % A.Hs); t /= A.Hs;
      const int od = (int)(t % A.Ds);
      n = (int)(t / A.Ds);
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
      for (int j = 0; j < TB; ++j) acc[i][j] += sv[i] * bv[j];
    }
  }
  __shared__ float red[PWG_T / 32][TS * TB + TS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < TS; ++i) {
#pragma unroll
    for (int j = 0; j < TB; ++j) {
      const float v = warp_sum(acc[i][j]);
      if (lane == 0) red[wid][i * TB + j] = v;
    }
    const float b = warp_sum(bs[i]);
    if (lane == 0) red[wid][TS * TB + i] = b;
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

template <bool SRC_IS_BIG>
static int launch_pointwise(const PwArgs& A, cudaStream_t st) {
  const long long nvox = (long long)A.N * A.Ds * A.Hs * A.Ws;
  const unsigned gx = (unsigned)((nvox + PW_T - 1) / PW_T);
  int co_t = A.Cout <= 4 ? 4 : A.Cout <= 8 ? 8 : A.Cout <= 12 ? 12 : 16;
  if (A.Cin * co_t > PW_MAX_W) return NAS3D_ERR_UNSUPPORTED;
  const unsigned gy = (unsigned)((A.Cout + co_t - 1) / co_t);
  switch (co_t) {
    case 4: pointwise_kernel<4, SRC_IS_BIG><<<dim3(gx, gy), PW_T, 0, st>>>(A); break;
    case 8: pointwise_kernel<8, SRC_IS_BIG><<<dim3(gx, gy), PW_T, 0, st>>>(A); break;
    case 12: pointwise_kernel<12, SRC_IS_BIG><<<dim3(gx, gy), PW_T, 0, st>>>(A); break;
    default: pointwise_kernel<16, SRC_IS_BIG><<<dim3(gx, gy), PW_T, 0, st>>>(A); break;
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
  if (moments && (((long long)d->Ds * d->Hs * d->Ws) % PW_T != 0 || accumulate || sigmoid))
    return NAS3D_ERR_UNSUPPORTED;
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
  const long long per_block = (long long)PWG_T * PWG_ITER;
  const unsigned gx = (unsigned)((nvox + per_block - 1) / per_block);
  const int Cs = d->Cs, Cb = d->Cb;
  const unsigned gy = (unsigned)((Cs + 3) / 4);
  if (Cb % 12 == 0 && Cb <= 48) {
    pointwise_wgrad_kernel<4, 12><<<dim3(gx, gy, Cb / 12), PWG_T, 0, st>>>(A, small, big, Cs, Cb, d->ld_small, d->ld_big, dW, dbias_small);
  } else if (Cb % 8 == 0) {
    pointwise_wgrad_kernel<4, 8><<<dim3(gx, gy, Cb / 8), PWG_T, 0, st>>>(A, small, big, Cs, Cb, d->ld_small, d->ld_big, dW, dbias_small);
  } else {
    pointwise_wgrad_kernel<4, 4><<<dim3(gx, gy, (Cb + 3) / 4), PWG_T, 0, st>>>(A, small, big, Cs, Cb, d->ld_small, d->ld_big, dW, dbias_small);
  }
  return launched("pointwise_wgrad");
}

}  // namespace nas3d
