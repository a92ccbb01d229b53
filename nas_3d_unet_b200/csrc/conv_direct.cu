// Generic direct (gather) 3-D convolution kernels on NDHWC fp32 — the shape-agnostic path
// behind nn.Conv3d / nn.ConvTranspose3d / convolution_backward (prim_ops.py:95-110,142-147).
// Covers every combination the network uses: k in {1,3}, stride in {1,2}, dilation in {1,2},
// groups in {1,C}, plain and transposed, forward / dgrad / wgrad.
//
// "Conv view" (see nas3d_b200.h): big_pos = small_pos*stride - pad + tap*dil,
// W[Cs][Cb][k^3].  Two gather directions share one kernel template:
//   SFB (small-from-big): Conv3d forward, ConvTranspose3d dgrad   (reduce over Cb)
//   BFS (big-from-small): ConvTranspose3d forward, Conv3d dgrad   (reduce over Cs)
// The hot dense 3x3x3 shapes are taken over by specialised kernels (conv_tiled.cu /
// conv_umma.cu); this file is the always-correct fallback and the odd-shape path.
#include "common.cuh"
#include "conv_tiled.h"

#include <stdlib.h>

namespace nas3d {

static bool tiled_enabled() { return g_opt.tiled != 0; }   // option: parity tests A/B the paths
static bool pointwise_shape(const nas3d_conv_desc* d) {
  return tiled_enabled() && d->k == 1 && !d->depthwise && d->pad == 0;
}
// 1: stride-1 (dil 1|2) ; 2: stride-2 dilation-2 (= stride-1 dil-1 on the even sub-lattice of big)
static int tiled_kind(const nas3d_conv_desc* d, bool wgrad = false) {
  if (!tiled_enabled() || d->k != 3 || d->depthwise || d->Cb != d->Cs) return 0;
  if (wgrad) {
    if (d->Cb % 4 || d->Cb > 64) return 0;
  } else if (!(d->Cb == 4 || d->Cb == 8 || d->Cb == 16)) {
    return 0;
  }
  if (d->stride == 1 && d->pad == d->dil && (d->dil == 1 || d->dil == 2)) return 1;
  if (d->stride == 2 && d->dil == 2 && d->pad == 2 && d->Db == 2 * d->Ds && d->Hb == 2 * d->Hs &&
      d->Wb == 2 * d->Ws)
    return 2;
  return 0;
}
static bool s2d1_shape(const nas3d_conv_desc* d) {
  return tiled_enabled() && d->k == 3 && d->stride == 2 && d->dil == 1 && d->pad == 1 && !d->depthwise;
}
static bool dw3_shape(const nas3d_conv_desc* d) {
  return tiled_enabled() && d->depthwise && d->k == 3 && d->dil == 1 && d->pad == 1 &&
         (d->stride == 1 || d->stride == 2) && d->Cb == d->Cs;
}
static S2Args s2_args(const nas3d_conv_desc* d, const float* big, const float* small, const float* w) {
  S2Args A{};
  A.big = const_cast<float*>(big); A.small = const_cast<float*>(small); A.w = w;
  A.N = d->N;
  A.Db = d->Db; A.Hb = d->Hb; A.Wb = d->Wb; A.Cb = d->Cb; A.ld_big = d->ld_big;
  A.Ds = d->Ds; A.Hs = d->Hs; A.Ws = d->Ws; A.Cs = d->Cs; A.ld_small = d->ld_small;
  return A;
}
static void tiled_fill(const nas3d_conv_desc* d, int kind, bool x_is_big, TiledArgs* T) {
  T->N = d->N; T->D = d->Ds; T->H = d->Hs; T->W = d->Ws;
  const int bs = kind == 2 ? 2 : 1;
  if (x_is_big) {
    T->xs = bs; T->Dx = d->Db; T->Hx = d->Hb; T->Wx = d->Wb; T->ldx = d->ld_big;
    T->ys = 1;  T->Dy = d->Ds; T->Hy = d->Hs; T->Wy = d->Ws; T->ldy = d->ld_small;
  } else {
    T->xs = 1;  T->Dx = d->Ds; T->Hx = d->Hs; T->Wx = d->Ws; T->ldx = d->ld_small;
    T->ys = bs; T->Dy = d->Db; T->Hy = d->Hb; T->Wy = d->Wb; T->ldy = d->ld_big;
  }
  T->tiles_w = T->tiles_h = T->tiles_d = 0;
}

struct ConvArgs {
  const float* src;
  const float* w;
  const float* bias;
  const float* scale;   // SFB: per-(n,cb) prologue scale; BFS: per-(n,cb) epilogue scale
  const float* mask;    // BFS: epilogue relu mask tensor (same geometry as big)
  float* dst;
  int N;
  int Db, Hb, Wb, Cb, ldb;
  int Ds, Hs, Ws, Cs, lds;
  int k, stride, dil, pad;
  int relu, sigmoid, accumulate, ld_mask;
};

constexpr int CT = 128;          // threads per conv block (one output voxel each)
constexpr int RED_CHUNK3 = 16;   // reduction-channel chunk for k=3
constexpr int RED_CHUNK1 = 64;   // reduction-channel chunk for k=1
constexpr int WSM_FLOATS = 27 * RED_CHUNK3 * 16;  // 6912 floats = 27 KB

// dst voxel <- gather over src.  PT = produced channels per thread.
template <bool BFS, int PT>
__global__ void __launch_bounds__(CT) conv_gather_kernel(const ConvArgs A) {
  __shared__ __align__(16) float Wsm[WSM_FLOATS];
  const int T = A.k * A.k * A.k;
  // geometry of produced (dst) and reduced (src) tensors
  const int Dd = BFS ? A.Db : A.Ds, Hd = BFS ? A.Hb : A.Hs, Wd = BFS ? A.Wb : A.Ws;
  const int Dr = BFS ? A.Ds : A.Db, Hr = BFS ? A.Hs : A.Hb, Wr = BFS ? A.Ws : A.Wb;
  const int Cprod = BFS ? A.Cb : A.Cs, Cred = BFS ? A.Cs : A.Cb;
  const int ldd = BFS ? A.ldb : A.lds, ldr = BFS ? A.lds : A.ldb;
  const int chunk = (A.k == 1) ? RED_CHUNK1 : RED_CHUNK3;
  const bool vec = (Cred % 4 == 0) && (ldr % 4 == 0);

  const long long nvox = (long long)A.N * Dd * Hd * Wd;
  const long long o = (long long)blockIdx.x * CT + threadIdx.x;
  const bool valid = o < nvox;
  int n = 0, od = 0, oh = 0, ow = 0;
  if (valid) {
    long long t = o;
    ow = (int)(t % Wd); t /= Wd;
    oh = (int)(t % Hd); t /= Hd;
    od = (int)(t % Dd); n = (int)(t / Dd);
  }
  const int p0 = blockIdx.y * PT;  // first produced channel of this block

  float acc[PT];
#pragma unroll
  for (int j = 0; j < PT; ++j) acc[j] = 0.f;

  for (int c0 = 0; c0 < Cred; c0 += chunk) {
    const int cn = min(chunk, Cred - c0);
    __syncthreads();
    // Wsm[tap][cr][PT]  <-  W[cs][cb][tap]
    for (int i = threadIdx.x; i < T * cn * PT; i += CT) {
      const int j = i % PT;
      const int cr = (i / PT) % cn;
      const int tap = i / (PT * cn);
      const int cp = p0 + j;
      float v = 0.f;
      if (cp < Cprod) {
        const long long gi = BFS ? ((long long)(c0 + cr) * A.Cb + cp) * T + tap
                                 : ((long long)cp * A.Cb + (c0 + cr)) * T + tap;
        v = __ldg(A.w + gi);
      }
      Wsm[i] = v;
    }
    __syncthreads();
    if (!valid) continue;
    for (int kd = 0; kd < A.k; ++kd) {
      int id;
      if (BFS) {
        int num = od + A.pad - kd * A.dil;
        if (num < 0 || num % A.stride) continue;
        id = num / A.stride;
      } else {
        id = od * A.stride - A.pad + kd * A.dil;
      }
      if (id < 0 || id >= Dr) continue;
      for (int kh = 0; kh < A.k; ++kh) {
        int ih;
        if (BFS) {
          int num = oh + A.pad - kh * A.dil;
          if (num < 0 || num % A.stride) continue;
          ih = num / A.stride;
        } else {
          ih = oh * A.stride - A.pad + kh * A.dil;
        }
        if (ih < 0 || ih >= Hr) continue;
        for (int kw = 0; kw < A.k; ++kw) {
          int iw;
          if (BFS) {
            int num = ow + A.pad - kw * A.dil;
            if (num < 0 || num % A.stride) continue;
            iw = num / A.stride;
          } else {
            iw = ow * A.stride - A.pad + kw * A.dil;
          }
          if (iw < 0 || iw >= Wr) continue;
          const int tap = (kd * A.k + kh) * A.k + kw;
          const float* px = A.src + ((((long long)n * Dr + id) * Hr + ih) * Wr + iw) * ldr + c0;
          const float* pw = Wsm + (long long)tap * cn * PT;
          for (int c4 = 0; c4 < cn; c4 += 4) {
            const int nv = min(4, cn - c4);
            float4 x4;
            if (vec) {
              x4 = ldg4(px + c4);
            } else {
              x4.x = __ldg(px + c4);
              x4.y = nv > 1 ? __ldg(px + c4 + 1) : 0.f;
              x4.z = nv > 2 ? __ldg(px + c4 + 2) : 0.f;
              x4.w = nv > 3 ? __ldg(px + c4 + 3) : 0.f;
            }
            if (!BFS) {
              if (A.relu) {
                x4.x = fmaxf(x4.x, 0.f); x4.y = fmaxf(x4.y, 0.f);
                x4.z = fmaxf(x4.z, 0.f); x4.w = fmaxf(x4.w, 0.f);
              }
              if (A.scale) {
                const float* ps = A.scale + (long long)n * A.Cb + c0 + c4;
                x4.x *= __ldg(ps);
                if (nv > 1) x4.y *= __ldg(ps + 1);
                if (nv > 2) x4.z *= __ldg(ps + 2);
                if (nv > 3) x4.w *= __ldg(ps + 3);
              }
            }
            const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (e >= nv) break;
              const float4* wr = reinterpret_cast<const float4*>(pw + (c4 + e) * PT);
#pragma unroll
              for (int j4 = 0; j4 < PT / 4; ++j4) {
                const float4 w4 = wr[j4];
                acc[j4 * 4 + 0] += xv[e] * w4.x;
                acc[j4 * 4 + 1] += xv[e] * w4.y;
                acc[j4 * 4 + 2] += xv[e] * w4.z;
                acc[j4 * 4 + 3] += xv[e] * w4.w;
              }
            }
          }
        }
      }
    }
  }
  if (!valid) return;

  float* pd = A.dst + o * ldd + p0;
#pragma unroll
  for (int j = 0; j < PT; ++j) {
    const int cp = p0 + j;
    if (cp >= Cprod) break;
    float v = acc[j];
    if (A.bias) v += __ldg(A.bias + cp);
    if (BFS) {
      if (A.mask) v = (__ldg(A.mask + o * A.ld_mask + cp) > 0.f) ? v : 0.f;
      if (A.scale) v *= __ldg(A.scale + (long long)n * A.Cb + cp);
    }
    if (A.sigmoid) v = 1.f / (1.f + expf(-v));
    if (A.accumulate) v += pd[j];
    pd[j] = v;
  }
}

// depthwise (groups = C): one thread per (dst voxel, 4-channel group). W[c][0][tap].
template <bool BFS>
__global__ void __launch_bounds__(256) dw_gather_kernel(const ConvArgs A) {
  extern __shared__ __align__(16) float Wd_sm[];  // [tap][C]
  const int T = A.k * A.k * A.k;
  const int C = A.Cb;
  for (int i = threadIdx.x; i < T * C; i += blockDim.x) {
    const int c = i % C, tap = i / C;
    Wd_sm[i] = __ldg(A.w + (long long)c * T + tap);
  }
  __syncthreads();
  const int Dd = BFS ? A.Db : A.Ds, Hd = BFS ? A.Hb : A.Hs, Wd = BFS ? A.Wb : A.Ws;
  const int Dr = BFS ? A.Ds : A.Db, Hr = BFS ? A.Hs : A.Hb, Wr = BFS ? A.Ws : A.Wb;
  const int ldd = BFS ? A.ldb : A.lds, ldr = BFS ? A.lds : A.ldb;
  const int C4 = C / 4;
  const long long total = (long long)A.N * Dd * Hd * Wd * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long o = i / C4;
    const int c = (int)(i - o * C4) * 4;
    long long t = o;
    const int ow = (int)(t % Wd); t /= Wd;
    const int oh = (int)(t % Hd); t /= Hd;
    const int od = (int)(t % Dd);
    const int n = (int)(t / Dd);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int kd = 0; kd < A.k; ++kd) {
      int id;
      if (BFS) { int num = od + A.pad - kd * A.dil; if (num < 0 || num % A.stride) continue; id = num / A.stride; }
      else id = od * A.stride - A.pad + kd * A.dil;
      if (id < 0 || id >= Dr) continue;
      for (int kh = 0; kh < A.k; ++kh) {
        int ih;
        if (BFS) { int num = oh + A.pad - kh * A.dil; if (num < 0 || num % A.stride) continue; ih = num / A.stride; }
        else ih = oh * A.stride - A.pad + kh * A.dil;
        if (ih < 0 || ih >= Hr) continue;
        for (int kw = 0; kw < A.k; ++kw) {
          int iw;
          if (BFS) { int num = ow + A.pad - kw * A.dil; if (num < 0 || num % A.stride) continue; iw = num / A.stride; }
          else iw = ow * A.stride - A.pad + kw * A.dil;
          if (iw < 0 || iw >= Wr) continue;
          const int tap = (kd * A.k + kh) * A.k + kw;
          const float4 x4 = ldg4(A.src + ((((long long)n * Dr + id) * Hr + ih) * Wr + iw) * ldr + c);
          const float4 w4 = *reinterpret_cast<const float4*>(Wd_sm + tap * C + c);
          acc.x += x4.x * w4.x; acc.y += x4.y * w4.y; acc.z += x4.z * w4.z; acc.w += x4.w * w4.w;
        }
      }
    }
    if (A.bias) {
      const float4 b4 = ldg4(A.bias + c);
      acc.x += b4.x; acc.y += b4.y; acc.z += b4.z; acc.w += b4.w;
    }
    float* pd = A.dst + o * ldd + c;
    if (A.accumulate) {
      const float4 o4 = *reinterpret_cast<const float4*>(pd);
      acc.x += o4.x; acc.y += o4.y; acc.z += o4.z; acc.w += o4.w;
    }
    st4(pd, acc);
  }
}

// ---------------------------------------------------------------------------------------
// wgrad (dense): dW[cs][cb][tap] += sum_o small[o,cs] * f(big[map(o,tap), cb])
// grid = (voxel chunks, taps, (Cs/4 tiles)*(Cb/4 tiles)); a thread owns a 4x4 (cs,cb) tile.
// ---------------------------------------------------------------------------------------
constexpr int WG_T = 256;
constexpr int WG_ITER = 16;

__device__ __forceinline__ void load_upto4(const float* p, int nvalid, bool vec, float v[4]) {
  if (vec) {
    const float4 t = ldg4(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = (e < nvalid) ? __ldg(p + e) : 0.f;
  }
}

__global__ void __launch_bounds__(WG_T)
    wgrad_dense_kernel(const ConvArgs A, const float* __restrict__ small,
                       const float* __restrict__ big, float* __restrict__ dW,
                       float* __restrict__ dbias_small) {
  const int T = A.k * A.k * A.k;
  const int tap = blockIdx.y;
  const int kw = tap % A.k, kh = (tap / A.k) % A.k, kd = tap / (A.k * A.k);
  const int tb = (A.Cb + 3) / 4;
  const int cs0 = (blockIdx.z / tb) * 4, cb0 = (blockIdx.z % tb) * 4;
  const int ns = min(4, A.Cs - cs0), nb = min(4, A.Cb - cb0);
  const bool vec_s = (A.lds % 4 == 0) && ns == 4, vec_b = (A.ldb % 4 == 0) && nb == 4;
  const long long nvox = (long long)A.N * A.Ds * A.Hs * A.Ws;

  float acc[4][4];
  float bs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool do_bias = dbias_small != nullptr && tap == 0 && cb0 == 0;

  for (int it = 0; it < WG_ITER; ++it) {
    const long long o = ((long long)blockIdx.x * WG_ITER + it) * WG_T + threadIdx.x;
    if (o >= nvox) break;
    long long t = o;
    const int ow = (int)(t % A.Ws); t /= A.Ws;
    const int oh = (int)(t % A.Hs); t /= A.Hs;
    const int od = (int)(t % A.Ds);
    const int n = (int)(t / A.Ds);
    float sv[4];
    const int id = od * A.stride - A.pad + kd * A.dil;
    const int ih = oh * A.stride - A.pad + kh * A.dil;
    const int iw = ow * A.stride - A.pad + kw * A.dil;
    const bool inb = id >= 0 && id < A.Db && ih >= 0 && ih < A.Hb && iw >= 0 && iw < A.Wb;
    if (!inb && !do_bias) continue;
    load_upto4(small + o * A.lds + cs0, ns, vec_s, sv);
    if (do_bias) {
#pragma unroll
      for (int e = 0; e < 4; ++e) bs[e] += sv[e];
    }
    if (!inb) continue;
    float bv[4];
    load_upto4(big + ((((long long)n * A.Db + id) * A.Hb + ih) * A.Wb + iw) * A.ldb + cb0, nb,
               vec_b, bv);
    if (A.relu) {
#pragma unroll
      for (int e = 0; e < 4; ++e) bv[e] = fmaxf(bv[e], 0.f);
    }
    if (A.scale) {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (e < nb) bv[e] *= __ldg(A.scale + (long long)n * A.Cb + cb0 + e);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += sv[i] * bv[j];
  }

  __shared__ float red[WG_T / 32][20];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = warp_sum(acc[i][j]);
      if (lane == 0) red[wid][i * 4 + j] = v;
    }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float v = warp_sum(bs[e]);
    if (lane == 0) red[wid][16 + e] = v;
  }
  __syncthreads();
  if (threadIdx.x < 20) {
    float v = 0.f;
    for (int wdx = 0; wdx < WG_T / 32; ++wdx) v += red[wdx][threadIdx.x];
    if (threadIdx.x < 16) {
      const int i = threadIdx.x / 4, j = threadIdx.x % 4;
      if (i < ns && j < nb && v != 0.f)
        atomicAdd(dW + ((long long)(cs0 + i) * A.Cb + (cb0 + j)) * T + tap, v);
    } else if (do_bias) {
      const int e = threadIdx.x - 16;
      if (e < ns) atomicAdd(dbias_small + cs0 + e, v);
    }
  }
}

// wgrad (depthwise): dW[c][tap] += sum_o small[o,c]*big[map(o,tap),c]; optional dbias_small.
__global__ void __launch_bounds__(WG_T)
    wgrad_dw_kernel(const ConvArgs A, const float* __restrict__ small,
                    const float* __restrict__ big, float* __restrict__ dW,
                    float* __restrict__ dbias_small) {
  extern __shared__ float sm_acc[];  // [C] + [C]
  const int C = A.Cb, C4 = C / 4;
  const int T = A.k * A.k * A.k;
  const int tap = blockIdx.y;
  const int kw = tap % A.k, kh = (tap / A.k) % A.k, kd = tap / (A.k * A.k);
  for (int i = threadIdx.x; i < 2 * C; i += WG_T) sm_acc[i] = 0.f;
  __syncthreads();
  const bool do_bias = dbias_small != nullptr && tap == 0;
  const int c = (threadIdx.x % C4) * 4;   // WG_T % C4 == 0 (host-checked): fixed per thread
  const int vper = WG_T / C4;             // voxels per block-iteration
  const int vsub = threadIdx.x / C4;
  const long long nvox = (long long)A.N * A.Ds * A.Hs * A.Ws;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), bs = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int it = 0; it < WG_ITER * 4; ++it) {
    const long long o = ((long long)blockIdx.x * WG_ITER * 4 + it) * vper + vsub;
    if (o >= nvox) break;
    long long t = o;
    const int ow = (int)(t % A.Ws); t /= A.Ws;
    const int oh = (int)(t % A.Hs); t /= A.Hs;
    const int od = (int)(t % A.Ds);
    const int n = (int)(t / A.Ds);
    const int id = od * A.stride - A.pad + kd * A.dil;
    const int ih = oh * A.stride - A.pad + kh * A.dil;
    const int iw = ow * A.stride - A.pad + kw * A.dil;
    const bool inb = id >= 0 && id < A.Db && ih >= 0 && ih < A.Hb && iw >= 0 && iw < A.Wb;
    if (!inb && !do_bias) continue;
    const float4 s4 = ldg4(small + o * A.lds + c);
    if (do_bias) { bs.x += s4.x; bs.y += s4.y; bs.z += s4.z; bs.w += s4.w; }
    if (!inb) continue;
    const float4 b4 = ldg4(big + ((((long long)n * A.Db + id) * A.Hb + ih) * A.Wb + iw) * A.ldb + c);
    acc.x += s4.x * b4.x; acc.y += s4.y * b4.y; acc.z += s4.z * b4.z; acc.w += s4.w * b4.w;
  }
  atomicAdd(&sm_acc[c + 0], acc.x); atomicAdd(&sm_acc[c + 1], acc.y);
  atomicAdd(&sm_acc[c + 2], acc.z); atomicAdd(&sm_acc[c + 3], acc.w);
  if (do_bias) {
    atomicAdd(&sm_acc[C + c + 0], bs.x); atomicAdd(&sm_acc[C + c + 1], bs.y);
    atomicAdd(&sm_acc[C + c + 2], bs.z); atomicAdd(&sm_acc[C + c + 3], bs.w);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += WG_T) {
    atomicAdd(dW + (long long)i * T + tap, sm_acc[i]);
    if (do_bias) atomicAdd(dbias_small + i, sm_acc[C + i]);
  }
}

// per-channel sum over all voxels: out[c] += sum x[.,c]  (bias gradient on the big side)
__global__ void __launch_bounds__(256)
    colsum_kernel(const float* __restrict__ x, long long nvox, int C, int ld,
                  float* __restrict__ out) {
  extern __shared__ float sm_cs[];  // [C]
  const int C4 = C / 4;
  const int nthr = (256 / C4) * C4;
  for (int i = threadIdx.x; i < C; i += 256) sm_cs[i] = 0.f;
  __syncthreads();
  if (threadIdx.x < nthr) {
    const int c = (threadIdx.x % C4) * 4;
    const int vper = nthr / C4, vsub = threadIdx.x / C4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int it = 0; it < 64; ++it) {
      const long long o = ((long long)blockIdx.x * 64 + it) * vper + vsub;
      if (o >= nvox) break;
      const float4 v = ldg4(x + o * ld + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    atomicAdd(&sm_cs[c + 0], acc.x); atomicAdd(&sm_cs[c + 1], acc.y);
    atomicAdd(&sm_cs[c + 2], acc.z); atomicAdd(&sm_cs[c + 3], acc.w);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 256) atomicAdd(out + i, sm_cs[i]);
}

static int fill_args(const nas3d_conv_desc* d, ConvArgs* A) {
  NAS3D_REQUIRE(d != nullptr, "conv: NULL descriptor");
  NAS3D_REQUIRE(d->k == 1 || d->k == 3, "conv: kernel size %d unsupported (1 or 3)", d->k);
  NAS3D_REQUIRE(d->stride >= 1 && d->dil >= 1 && d->pad >= 0, "conv: bad stride/dil/pad");
  NAS3D_REQUIRE(d->N > 0 && d->Cb > 0 && d->Cs > 0, "conv: empty tensor");
  for (int ax = 0; ax < 3; ++ax) {
    const int b = ax == 0 ? d->Db : ax == 1 ? d->Hb : d->Wb;
    const int s = ax == 0 ? d->Ds : ax == 1 ? d->Hs : d->Ws;
    // the small extent must be what conv arithmetic gives for the big extent
    const int expect = (b + 2 * d->pad - d->dil * (d->k - 1) - 1) / d->stride + 1;
    NAS3D_REQUIRE(s == expect, "conv: small extent %d != %d implied by big extent %d", s, expect, b);
  }
  if (d->depthwise) NAS3D_REQUIRE(d->Cb == d->Cs, "conv: depthwise needs Cb == Cs");
  A->N = d->N;
  A->Db = d->Db; A->Hb = d->Hb; A->Wb = d->Wb; A->Cb = d->Cb; A->ldb = d->ld_big;
  A->Ds = d->Ds; A->Hs = d->Hs; A->Ws = d->Ws; A->Cs = d->Cs; A->lds = d->ld_small;
  A->k = d->k; A->stride = d->stride; A->dil = d->dil; A->pad = d->pad;
  A->relu = 0; A->sigmoid = 0; A->accumulate = 0; A->ld_mask = 0;
  A->src = nullptr; A->w = nullptr; A->bias = nullptr; A->scale = nullptr; A->mask = nullptr;
  A->dst = nullptr;
  return NAS3D_OK;
}

template <bool BFS>
static int launch_gather(const ConvArgs& A, bool depthwise, cudaStream_t st) {
  const int Cprod = BFS ? A.Cb : A.Cs, Cred = BFS ? A.Cs : A.Cb;
  const int ldr = BFS ? A.lds : A.ldb;
  const long long nvox = BFS ? (long long)A.N * A.Db * A.Hb * A.Wb
                             : (long long)A.N * A.Ds * A.Hs * A.Ws;
  NAS3D_REQUIRE((Cred % 4 != 0 || ldr % 4 != 0) || aligned16(A.src), "conv: src not 16B aligned");
  if (depthwise) {
    NAS3D_REQUIRE(Cred % 4 == 0 && ldr % 4 == 0, "depthwise conv: C=%d pitch %d", Cred, ldr);
    const int ldd = BFS ? A.ldb : A.lds;
    NAS3D_REQUIRE(ldd % 4 == 0 && aligned16(A.dst), "depthwise conv: dst misaligned");
    NAS3D_REQUIRE(!A.relu && !A.scale && !A.mask && !A.sigmoid, "depthwise conv: no prologue/epilogue");
    const int T = A.k * A.k * A.k;
    const long long total = nvox * (A.Cb / 4);
    long long blocks = (total + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    dw_gather_kernel<BFS><<<(unsigned)blocks, 256, sizeof(float) * T * A.Cb, st>>>(A);
    return launched("dw_gather");
  }
  const unsigned gx = (unsigned)((nvox + CT - 1) / CT);
  if (Cprod % 16 == 0) {
    conv_gather_kernel<BFS, 16><<<dim3(gx, Cprod / 16), CT, 0, st>>>(A);
  } else if (Cprod % 8 == 0) {
    conv_gather_kernel<BFS, 8><<<dim3(gx, Cprod / 8), CT, 0, st>>>(A);
  } else {
    conv_gather_kernel<BFS, 4><<<dim3(gx, (Cprod + 3) / 4), CT, 0, st>>>(A);
  }
  return launched("conv_gather");
}

}  // namespace nas3d

using namespace nas3d;

extern "C" {

// moments of a freshly written produced tensor for the paths that cannot fuse them
static int moments_fallback(double* moments, const float* y, int N, long long V, int C, int ld,
                            void* stream) {
  if (!moments) return NAS3D_OK;
  return nas3d_moments_nc(y, N, V, C, ld, moments, stream);
}

int nas3d_conv_small_from_big(const nas3d_conv_desc* d, const float* big, const float* w,
                              const float* bias, const float* big_scale, int big_relu,
                              int out_sigmoid, float* small, int accumulate, double* moments,
                              void* stream) {
  const long long Vs = d ? (long long)d->Ds * d->Hs * d->Ws : 0;
  if (moments && d) NAS3D_CUDA(cudaMemsetAsync(moments, 0, sizeof(double) * 2 * (size_t)d->N * d->Cs, (cudaStream_t)stream));
  ConvArgs A;
  int rc = fill_args(d, &A);
  if (rc) return rc;
  A.src = big; A.w = w; A.bias = bias; A.scale = big_scale; A.relu = big_relu;
  A.sigmoid = out_sigmoid; A.dst = small; A.accumulate = accumulate;
  if (pointwise_shape(d)) {
    // GroupNorm statistics are fused when the tiles do not straddle samples, else a separate pass
    rc = pointwise_sfb(d, big, w, bias, big_scale, big_relu, out_sigmoid, small, accumulate,
                       moments, (cudaStream_t)stream);
    if (rc != NAS3D_ERR_UNSUPPORTED) return rc;
    if (moments) {
      rc = pointwise_sfb(d, big, w, bias, big_scale, big_relu, out_sigmoid, small, accumulate,
                         nullptr, (cudaStream_t)stream);
      if (rc != NAS3D_ERR_UNSUPPORTED)
        return rc ? rc : moments_fallback(moments, small, d->N, Vs, d->Cs, d->ld_small, stream);
    }
  }
  if (dw3_shape(d) && !big_scale && !big_relu && !out_sigmoid) {
    if (d->stride == 1) {
      TiledArgs T{};
      T.x = big; T.w = w; T.bias = bias; T.y = small; T.accumulate = accumulate;
      T.N = d->N; T.D = d->Ds; T.H = d->Hs; T.W = d->Ws; T.ldx = d->ld_big; T.ldy = d->ld_small;
      rc = tiled_dw_s1(false, T, d->Cb, (cudaStream_t)stream);
    } else {
      S2Args S = s2_args(d, big, small, w);
      S.bias = bias; S.accumulate = accumulate;
      rc = tiled_dw_s2_sfb(S, (cudaStream_t)stream);
    }
    if (rc != NAS3D_ERR_UNSUPPORTED)
      return rc ? rc : moments_fallback(moments, small, d->N, Vs, d->Cs, d->ld_small, stream);
  }
  if (s2d1_shape(d) && !big_scale && !big_relu && !out_sigmoid) {
    S2Args S = s2_args(d, big, small, w);
    S.bias = bias; S.accumulate = accumulate; S.moments = moments;
    rc = tiled_s2_sfb(S, (cudaStream_t)stream);
    if (rc != NAS3D_ERR_UNSUPPORTED) return rc;
  }
  if (const int kind = tiled_kind(d); kind && !big_scale && !big_relu && !out_sigmoid) {
    TiledArgs T{};
    tiled_fill(d, kind, true, &T);
    T.x = big; T.w = w; T.bias = bias; T.y = small; T.accumulate = accumulate; T.moments = moments;
    rc = tiled_conv3_s1(d->Cb, kind == 2 ? 1 : d->dil, false, T, (cudaStream_t)stream);
    if (rc != NAS3D_ERR_UNSUPPORTED) return rc;
  }
  rc = launch_gather<false>(A, d->depthwise != 0, (cudaStream_t)stream);
  return rc ? rc : moments_fallback(moments, small, d->N, Vs, d->Cs, d->ld_small, stream);
}

int nas3d_conv_big_from_small(const nas3d_conv_desc* d, const float* small, const float* w,
                              const float* bias, const float* mask_big, int ld_mask,
                              const float* big_scale, float* big, int accumulate, double* moments,
                              void* stream) {
  const long long Vb = d ? (long long)d->Db * d->Hb * d->Wb : 0;
  if (moments && d) NAS3D_CUDA(cudaMemsetAsync(moments, 0, sizeof(double) * 2 * (size_t)d->N * d->Cb, (cudaStream_t)stream));
  ConvArgs A;
  int rc = fill_args(d, &A);
  if (rc) return rc;
  A.src = small; A.w = w; A.bias = bias; A.mask = mask_big; A.ld_mask = ld_mask;
  A.scale = big_scale; A.dst = big; A.accumulate = accumulate;
  if (pointwise_shape(d)) {
    rc = pointwise_bfs(d, small, w, bias, mask_big, ld_mask, big_scale, big, accumulate,
                       (cudaStream_t)stream);
    if (rc != NAS3D_ERR_UNSUPPORTED)
      return rc ? rc : moments_fallback(moments, big, d->N, Vb, d->Cb, d->ld_big, stream);
  }
  if (dw3_shape(d) && !big_scale && !mask_big) {
    if (d->stride == 1) {
      TiledArgs T{};
      T.x = small; T.w = w; T.bias = bias; T.y = big; T.accumulate = accumulate;
      T.N = d->N; T.D = d->Db; T.H = d->Hb; T.W = d->Wb; T.ldx = d->ld_small; T.ldy = d->ld_big;
      rc = tiled_dw_s1(true, T, d->Cb, (cudaStream_t)stream);
    } else {
      S2Args S = s2_args(d, big, small, w);
      S.bias = bias; S.accumulate = accumulate;
      rc = tiled_dw_s2_bfs(S, (cudaStream_t)stream);
    }
    if (rc != NAS3D_ERR_UNSUPPORTED)
      return rc ? rc : moments_fallback(moments, big, d->N, Vb, d->Cb, d->ld_big, stream);
  }
  if (s2d1_shape(d) && !big_scale && !mask_big) {
    S2Args S = s2_args(d, big, small, w);
    S.bias = bias; S.accumulate = accumulate; S.moments = moments;
    rc = tiled_s2_bfs(S, (cudaStream_t)stream);
    if (rc != NAS3D_ERR_UNSUPPORTED) return rc;
  }
  if (const int kind = tiled_kind(d);
      kind && !big_scale && !mask_big && (kind == 1 || accumulate || d->ld_big == d->Cb) &&
      d->Ws >= 8 && d->ld_small % 4 == 0 && d->ld_big % 4 == 0 && aligned16(small) && aligned16(big)) {
    cudaStream_t st = (cudaStream_t)stream;
    if (kind == 2 && !accumulate) {
      // voxels off the even lattice receive no tap: they are bias (transposed fwd) or 0 (dgrad)
      const long long nvb = (long long)d->N * d->Db * d->Hb * d->Wb;
      if (bias) {
        rc = fill_channels(big, bias, nvb, d->Cb, d->ld_big, st);
        if (rc) return rc;
      } else {
        NAS3D_CUDA(cudaMemsetAsync(big, 0, sizeof(float) * (size_t)nvb * d->Cb, st));
      }
    }
    TiledArgs T{};
    tiled_fill(d, kind, false, &T);
    T.x = small; T.w = w; T.bias = bias; T.y = big; T.accumulate = accumulate;
    T.moments = kind == 1 ? moments : nullptr;    // lattice writes skip the bias-only voxels
    rc = tiled_conv3_s1(d->Cb, kind == 2 ? 1 : d->dil, true, T, st);
    if (rc != NAS3D_ERR_UNSUPPORTED)
      return (rc || kind == 1) ? rc : moments_fallback(moments, big, d->N, Vb, d->Cb, d->ld_big, stream);
  }
  rc = launch_gather<true>(A, d->depthwise != 0, (cudaStream_t)stream);
  return rc ? rc : moments_fallback(moments, big, d->N, Vb, d->Cb, d->ld_big, stream);
}

long long nas3d_conv_wgrad_workspace_floats(const nas3d_conv_desc* d, int has_prologue) {
  if (!d || has_prologue || !tiled_enabled()) return 0;
  return umma_wgrad_workspace_floats(d);
}

int nas3d_conv_wgrad(const nas3d_conv_desc* d, const float* small, const float* big,
                     const float* big_scale, int big_relu, float* dW, float* d_bias_small,
                     float* d_bias_big, void* stream) {
  return nas3d_conv_wgrad_ws(d, small, big, big_scale, big_relu, dW, d_bias_small, d_bias_big, nullptr, 0,
                             stream);
}

int nas3d_conv_wgrad_ws(const nas3d_conv_desc* d, const float* small, const float* big,
                        const float* big_scale, int big_relu, float* dW, float* d_bias_small,
                        float* d_bias_big, float* workspace, long long workspace_floats, void* stream) {
  ConvArgs A;
  int rc = fill_args(d, &A);
  if (rc) return rc;
  A.scale = big_scale; A.relu = big_relu;
  cudaStream_t st = (cudaStream_t)stream;
  const int T = A.k * A.k * A.k;
  const long long nvox = (long long)A.N * A.Ds * A.Hs * A.Ws;
  bool done = false;
  if (pointwise_shape(d)) {
    rc = pointwise_wgrad(d, small, big, big_scale, big_relu, dW, d_bias_small, st);
    if (rc == NAS3D_OK) done = true;
    else if (rc != NAS3D_ERR_UNSUPPORTED) return rc;
    rc = NAS3D_OK;
  }
  if (!done && tiled_enabled() && !big_scale && !big_relu) {
    rc = umma_wgrad(d, small, big, dW, workspace, workspace_floats, st);   // wide dense 3x3x3: tcgen05 split-K GEMM
    if (rc == NAS3D_OK) {
      done = true;
      if (d_bias_small) {                          // not followed by a GroupNorm: column sums of dy
        const int C4 = A.Cs / 4;
        const long long per_block = 64LL * (256 / C4);
        colsum_kernel<<<(unsigned)((nvox + per_block - 1) / per_block), 256, sizeof(float) * A.Cs, st>>>(
            small, nvox, A.Cs, A.lds, d_bias_small);
        rc = launched("colsum");
        if (rc) return rc;
      }
    } else if (rc != NAS3D_ERR_UNSUPPORTED) return rc;
    rc = NAS3D_OK;
  }
  if (!done && dw3_shape(d) && !big_scale && !big_relu) {
    S2Args S = s2_args(d, big, small, nullptr);
    S.dW = dW; S.dbias_small = d_bias_small;
    rc = tiled_dw_wgrad(S, d->stride, st);
    if (rc == NAS3D_OK) done = true;
    else if (rc != NAS3D_ERR_UNSUPPORTED) return rc;
    rc = NAS3D_OK;
  }
  if (!done && s2d1_shape(d) && !big_scale && !big_relu) {
    S2Args S = s2_args(d, big, small, nullptr);
    S.dW = dW; S.dbias_small = d_bias_small;
    rc = tiled_s2_wgrad(S, st);
    if (rc == NAS3D_OK) done = true;
    else if (rc != NAS3D_ERR_UNSUPPORTED) return rc;
    rc = NAS3D_OK;
  }
  if (const int kind = tiled_kind(d, true); !done && kind && !big_scale && !big_relu) {
    WgradArgs T{};
    T.x = big; T.dy = small; T.dW = dW; T.dbias = d_bias_small;
    T.N = d->N; T.D = d->Ds; T.H = d->Hs; T.W = d->Ws; T.ldx = d->ld_big; T.ldy = d->ld_small;
    T.xs = kind == 2 ? 2 : 1; T.Dx = d->Db; T.Hx = d->Hb; T.Wx = d->Wb;
    T.ys = 1; T.Dy = d->Ds; T.Hy = d->Hs; T.Wy = d->Ws;
    rc = tiled_wgrad3_s1(d->Cb, kind == 2 ? 1 : d->dil, T, st);
    if (rc == NAS3D_OK) done = true;
    else if (rc != NAS3D_ERR_UNSUPPORTED) return rc;
    rc = NAS3D_OK;
  }
  if (done) {
  } else if (d->depthwise) {
    const int C4 = A.Cb / 4;
    NAS3D_REQUIRE(A.Cb % 4 == 0 && WG_T % C4 == 0 && A.lds % 4 == 0 && A.ldb % 4 == 0,
                  "depthwise wgrad: C=%d unsupported", A.Cb);
    NAS3D_REQUIRE(!big_relu && !big_scale, "depthwise wgrad: no prologue");
    const long long per_block = (long long)WG_ITER * 4 * (WG_T / C4);
    dim3 grid((unsigned)((nvox + per_block - 1) / per_block), T);
    wgrad_dw_kernel<<<grid, WG_T, sizeof(float) * 2 * A.Cb, st>>>(A, small, big, dW, d_bias_small);
    rc = launched("wgrad_dw");
  } else {
    const long long per_block = (long long)WG_ITER * WG_T;
    dim3 grid((unsigned)((nvox + per_block - 1) / per_block), T,
              ((A.Cs + 3) / 4) * ((A.Cb + 3) / 4));
    wgrad_dense_kernel<<<grid, WG_T, 0, st>>>(A, small, big, dW, d_bias_small);
    rc = launched("wgrad_dense");
  }
  if (rc) return rc;
  if (d_bias_big) {
    NAS3D_REQUIRE(A.Cb % 4 == 0 && A.ldb % 4 == 0 && A.Cb <= 1024, "d_bias_big: C=%d unsupported", A.Cb);
    const long long nvb = (long long)A.N * A.Db * A.Hb * A.Wb;
    const int C4 = A.Cb / 4;
    const long long per_block = 64LL * (256 / C4);
    colsum_kernel<<<(unsigned)((nvb + per_block - 1) / per_block), 256, sizeof(float) * A.Cb, st>>>(
        big, nvb, A.Cb, A.ldb, d_bias_big);
    rc = launched("colsum");
  }
  return rc;
}


// ---- virtual-concat 1x1x1 convolutions (cell.py:82 without materialising the concat) ----------
static int cat_check(const nas3d_conv_desc* d, int nparts) {
  NAS3D_REQUIRE(d && d->k == 1 && d->pad == 0 && !d->depthwise, "conv1x1_cat: needs a dense 1x1x1 conv");
  NAS3D_REQUIRE(nparts >= 1 && nparts <= 4 && d->Cb % nparts == 0 && (d->Cb / nparts) % 4 == 0,
                "conv1x1_cat: %d parts of %d channels unsupported", nparts, d->Cb);
  for (int ax = 0; ax < 3; ++ax) {
    const int b = ax == 0 ? d->Db : ax == 1 ? d->Hb : d->Wb;
    const int s = ax == 0 ? d->Ds : ax == 1 ? d->Hs : d->Ws;
    NAS3D_REQUIRE(s == (b - 1) / d->stride + 1, "conv1x1_cat: extents %d / %d", b, s);
  }
  return NAS3D_OK;
}

int nas3d_conv1x1_cat_fwd(const nas3d_conv_desc* d, int nparts, const float* const* big_parts,
                          const int* part_ld, const float* w, const float* bias,
                          const float* big_scale, int big_relu, int out_sigmoid, float* small,
                          double* moments, void* stream) {
  int rc = cat_check(d, nparts);
  if (rc) return rc;
  PwCat cat{nparts, big_parts, nullptr, nullptr, part_ld, nullptr, nullptr};
  if (moments)
    NAS3D_CUDA(cudaMemsetAsync(moments, 0, sizeof(double) * 2 * (size_t)d->N * d->Cs, (cudaStream_t)stream));
  rc = pointwise_sfb(d, nullptr, w, bias, big_scale, big_relu, out_sigmoid, small, 0, moments,
                     (cudaStream_t)stream, &cat);
  if (rc != NAS3D_ERR_UNSUPPORTED || !moments) {
    NAS3D_REQUIRE(rc != NAS3D_ERR_UNSUPPORTED, "conv1x1_cat_fwd: unsupported layout");
    return rc;
  }
  rc = pointwise_sfb(d, nullptr, w, bias, big_scale, big_relu, out_sigmoid, small, 0, nullptr,
                     (cudaStream_t)stream, &cat);
  NAS3D_REQUIRE(rc != NAS3D_ERR_UNSUPPORTED, "conv1x1_cat_fwd: unsupported layout");
  if (rc) return rc;
  return moments_fallback(moments, small, d->N, (long long)d->Ds * d->Hs * d->Ws, d->Cs, d->ld_small, stream);
}

int nas3d_conv1x1_cat_dgrad(const nas3d_conv_desc* d, int nparts, float* const* dbig_parts,
                            const int* part_ld, const int* accumulate, const float* small,
                            const float* w, const float* const* mask_parts, const int* mask_ld,
                            const float* big_scale, void* stream) {
  int rc = cat_check(d, nparts);
  if (rc) return rc;
  PwCat cat{nparts, nullptr, dbig_parts, mask_parts, part_ld, mask_ld, accumulate};
  rc = pointwise_bfs(d, small, w, nullptr, nullptr, 0, big_scale, nullptr, 0, (cudaStream_t)stream, &cat);
  NAS3D_REQUIRE(rc != NAS3D_ERR_UNSUPPORTED, "conv1x1_cat_dgrad: unsupported layout");
  return rc;
}

int nas3d_conv1x1_cat_wgrad(const nas3d_conv_desc* d, int nparts, const float* const* big_parts,
                            const int* part_ld, const float* small, const float* big_scale,
                            int big_relu, float* dW, float* d_bias_small, void* stream) {
  int rc = cat_check(d, nparts);
  if (rc) return rc;
  PwCat cat{nparts, big_parts, nullptr, nullptr, part_ld, nullptr, nullptr};
  rc = pointwise_wgrad(d, small, nullptr, big_scale, big_relu, dW, d_bias_small, (cudaStream_t)stream, &cat);
  NAS3D_REQUIRE(rc != NAS3D_ERR_UNSUPPORTED, "conv1x1_cat_wgrad: unsupported layout");
  return rc;
}

}  // extern "C"
