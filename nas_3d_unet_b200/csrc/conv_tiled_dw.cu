// Tiled DEPTHWISE 3x3x3 convolutions (groups = C, dilation 1): the depth_conv half of dep_conv /
// down_dep_conv / up_dep_conv (prim_ops.py:94-98,104-107), forward, dgrad and wgrad, stride 1 and 2.
// AI = 6.75 FLOP/B: HBM-bound streaming kernels.  Same tiling as conv_tiled.cu / conv_tiled_s2.cu,
// one 4-channel chunk per CTA (blockIdx.y), the per-channel weights W[c][0][27] held as 27 float4.
#include "common.cuh"
#include "conv_tiled.h"

namespace nas3d {

__device__ __forceinline__ void dcp16(void* smem_dst, const void* gsrc, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void dcp_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}
__device__ __forceinline__ void fma4(float4& a, const float4& x, const float4& w) {
  // two packed FFMA2 (common.cuh) instead of four FFMA
  float2 lo = make_float2(a.x, a.y), hi = make_float2(a.z, a.w);
  fma2v(lo, x.x, x.y, w.x, w.y);
  fma2v(hi, x.z, x.w, w.z, w.w);
  a = make_float4(lo.x, lo.y, hi.x, hi.y);
}
// stage wsm[tap] = {W[c0+0][tap], .., W[c0+3][tap]} (optionally mirrored taps)
__device__ __forceinline__ void stage_dw_weights(float4* wsm, const float* w, int c0, bool flip) {
  if (threadIdx.x < 27) {
    const int t = flip ? 26 - threadIdx.x : threadIdx.x;
    wsm[threadIdx.x] = make_float4(__ldg(w + (c0 + 0) * 27 + t), __ldg(w + (c0 + 1) * 27 + t),
                                   __ldg(w + (c0 + 2) * 27 + t), __ldg(w + (c0 + 3) * 27 + t));
  }
}

// =========================================================================================
// stride 1 (forward / dgrad with FLIP): tile 4 x 16 x 32, 256 threads, 2x4 voxels per thread
// =========================================================================================
constexpr int D1_TW = 32, D1_TH = 16, D1_TD = 4, D1_PW = 34, D1_PH = 18, D1_PD = 6;
constexpr int D1_THREADS = 256;
constexpr size_t D1_SMEM = sizeof(float4) * (D1_PD * D1_PH * D1_PW + 32);

template <bool FLIP>
__global__ void __launch_bounds__(D1_THREADS) dw3_s1_kernel(const TiledArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* tile = reinterpret_cast<float4*>(smem_raw);
  float4* wsm = tile + D1_PD * D1_PH * D1_PW;
  const int c0 = blockIdx.y * 4;
  int b = blockIdx.x;
  const int tw = b % A.tiles_w; b /= A.tiles_w;
  const int th = b % A.tiles_h; b /= A.tiles_h;
  const int td = b % A.tiles_d;
  const int n = b / A.tiles_d;
  const int w0 = tw * D1_TW, h0 = th * D1_TH, d0 = td * D1_TD;
  stage_dw_weights(wsm, A.w, c0, FLIP);
  const float* xb = A.x + (long long)n * A.D * A.H * A.W * A.ldx + c0;
  for (int i = threadIdx.x; i < D1_PD * D1_PH * D1_PW; i += D1_THREADS) {
    int r = i;
    const int pw = r % D1_PW; r /= D1_PW;
    const int ph = r % D1_PH;
    const int pd = r / D1_PH;
    const int gd = d0 - 1 + pd, gh = h0 - 1 + ph, gw = w0 - 1 + pw;
    const bool ok = gd >= 0 && gd < A.D && gh >= 0 && gh < A.H && gw >= 0 && gw < A.W;
    // 32-bit in-sample offset (host guarantees samples < 2^31 floats)
    const unsigned off = (unsigned)(((gd * A.H + gh) * A.W + gw) * A.ldx);
    dcp16(&tile[i], ok ? xb + off : A.x, ok);
  }
  dcp_wait_all();
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int hg = ty & 3, dg = ty >> 2;
  float4 acc[2][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int h = 0; h < 4; ++h) acc[a][h] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int kd = 0; kd < 3; ++kd)
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const float4 w0v = wsm[(kd * 3 + 0) * 3 + kw], w1v = wsm[(kd * 3 + 1) * 3 + kw],
                   w2v = wsm[(kd * 3 + 2) * 3 + kw];
#pragma unroll
      for (int od = 0; od < 2; ++od) {
        const float4* col = tile + ((dg * 2 + od + kd) * D1_PH + hg * 4) * D1_PW + tx + kw;
        float4 xin[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) xin[r] = col[r * D1_PW];
#pragma unroll
        for (int oh = 0; oh < 4; ++oh) {
          fma4(acc[od][oh], xin[oh], w0v);
          fma4(acc[od][oh], xin[oh + 1], w1v);
          fma4(acc[od][oh], xin[oh + 2], w2v);
        }
      }
    }
  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (A.bias) bias4 = ldg4(A.bias + c0);
  const int gw = w0 + tx;
  if (gw >= A.W) return;
  float* yb = A.y + (long long)n * A.D * A.H * A.W * A.ldy + c0;
#pragma unroll
  for (int od = 0; od < 2; ++od) {
    const int gd = d0 + dg * 2 + od;
    if (gd >= A.D) break;
#pragma unroll
    for (int oh = 0; oh < 4; ++oh) {
      const int gh = h0 + hg * 4 + oh;
      if (gh >= A.H) break;
      float* p = yb + (((long long)gd * A.H + gh) * A.W + gw) * A.ldy;
      float4 v = make_float4(acc[od][oh].x + bias4.x, acc[od][oh].y + bias4.y,
                             acc[od][oh].z + bias4.z, acc[od][oh].w + bias4.w);
      if (A.accumulate) {
        const float4 o = *reinterpret_cast<const float4*>(p);
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
      }
      st4(p, v);
    }
  }
}

// =========================================================================================
// stride 2, small-from-big: tile 2 x 8 x 32 small voxels, 128 threads
// =========================================================================================
constexpr int DS_TW = 32, DS_TH = 8, DS_TD = 2, DS_PD = 5, DS_PH = 17, DS_PWS = 65, DS_EVEN = 33;
constexpr int DS_THREADS = 128;
constexpr size_t DS_SMEM = sizeof(float4) * (DS_PD * DS_PH * DS_PWS + 32);

__global__ void __launch_bounds__(DS_THREADS) dw3_s2_sfb_kernel(const S2Args A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* tile = reinterpret_cast<float4*>(smem_raw);
  float4* wsm = tile + DS_PD * DS_PH * DS_PWS;
  const int c0 = blockIdx.y * 4;
  int b = blockIdx.x;
  const int tw = b % A.tiles_w; b /= A.tiles_w;
  const int th = b % A.tiles_h; b /= A.tiles_h;
  const int td = b % A.tiles_d;
  const int n = b / A.tiles_d;
  const int w0 = tw * DS_TW, h0 = th * DS_TH, d0 = td * DS_TD;
  stage_dw_weights(wsm, A.w, c0, false);
  const float* xb = A.big + (long long)n * A.Db * A.Hb * A.Wb * A.ld_big + c0;
  for (int i = threadIdx.x; i < DS_PD * DS_PH * DS_PWS; i += DS_THREADS) {
    int r = i;
    const int pwi = r % DS_PWS; r /= DS_PWS;
    const int ph = r % DS_PH;
    const int pd = r / DS_PH;
    const int bw = pwi < DS_EVEN ? 2 * pwi : 2 * (pwi - DS_EVEN) + 1;
    const int gd = 2 * d0 - 1 + pd, gh = 2 * h0 - 1 + ph, gw = 2 * w0 - 1 + bw;
    const bool ok = gd >= 0 && gd < A.Db && gh >= 0 && gh < A.Hb && gw >= 0 && gw < A.Wb;
    const unsigned off = (unsigned)(((gd * A.Hb + gh) * A.Wb + gw) * A.ld_big);
    dcp16(&tile[i], ok ? xb + off : A.big, ok);
  }
  dcp_wait_all();
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int hg = ty & 1, dg = ty >> 1;
  float4 acc[4];
#pragma unroll
  for (int h = 0; h < 4; ++h) acc[h] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int kd = 0; kd < 3; ++kd)
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int woff = (kw == 1) ? DS_EVEN + tx : tx + (kw >> 1);
      const float4* col = tile + ((2 * dg + kd) * DS_PH + 2 * (hg * 4)) * DS_PWS + woff;
      float4 xin[9];
#pragma unroll
      for (int r = 0; r < 9; ++r) xin[r] = col[r * DS_PWS];
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const float4 wv = wsm[(kd * 3 + kh) * 3 + kw];
#pragma unroll
        for (int oh = 0; oh < 4; ++oh) fma4(acc[oh], xin[2 * oh + kh], wv);
      }
    }
  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (A.bias) bias4 = ldg4(A.bias + c0);
  const int gw = w0 + tx, gd = d0 + dg;
  if (gw >= A.Ws || gd >= A.Ds) return;
  float* yb = A.small + (long long)n * A.Ds * A.Hs * A.Ws * A.ld_small + c0;
#pragma unroll
  for (int oh = 0; oh < 4; ++oh) {
    const int gh = h0 + hg * 4 + oh;
    if (gh >= A.Hs) break;
    float* p = yb + (((long long)gd * A.Hs + gh) * A.Ws + gw) * A.ld_small;
    float4 v = make_float4(acc[oh].x + bias4.x, acc[oh].y + bias4.y, acc[oh].z + bias4.z,
                           acc[oh].w + bias4.w);
    if (A.accumulate) {
      const float4 o = *reinterpret_cast<const float4*>(p);
      v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
    }
    st4(p, v);
  }
}

// =========================================================================================
// stride 2, big-from-small: big tile 4 x 16 x 64, 512 threads
// =========================================================================================
constexpr int DB_TDB = 4, DB_THB = 16, DB_TWB = 64, DB_SD = 3, DB_SH = 9, DB_SW = 33;
constexpr int DB_THREADS = 512;
constexpr size_t DB_SMEM = sizeof(float4) * (DB_SD * DB_SH * DB_SW + 32);

// register cap for two resident CTAs (78 registers x 512 threads left one): 130 -> 94 us at C4 8 x 64^3 -> 128^3
__global__ void __launch_bounds__(DB_THREADS, 2) dw3_s2_bfs_kernel(const S2Args A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* tile = reinterpret_cast<float4*>(smem_raw);
  float4* wsm = tile + DB_SD * DB_SH * DB_SW;
  const int c0 = blockIdx.y * 4;
  int b = blockIdx.x;
  const int tw = b % A.tiles_w; b /= A.tiles_w;
  const int th = b % A.tiles_h; b /= A.tiles_h;
  const int td = b % A.tiles_d;
  const int n = b / A.tiles_d;
  const int w0 = tw * DB_TWB, h0 = th * DB_THB, d0 = td * DB_TDB;
  stage_dw_weights(wsm, A.w, c0, false);
  const float* sb = A.small + (long long)n * A.Ds * A.Hs * A.Ws * A.ld_small + c0;
  for (int i = threadIdx.x; i < DB_SD * DB_SH * DB_SW; i += DB_THREADS) {
    int r = i;
    const int pw = r % DB_SW; r /= DB_SW;
    const int ph = r % DB_SH;
    const int pd = r / DB_SH;
    const int gd = d0 / 2 + pd, gh = h0 / 2 + ph, gw = w0 / 2 + pw;
    const bool ok = gd < A.Ds && gh < A.Hs && gw < A.Ws;
    const unsigned off = (unsigned)(((gd * A.Hs + gh) * A.Ws + gw) * A.ld_small);
    dcp16(&tile[i], ok ? sb + off : A.small, ok);
  }
  dcp_wait_all();
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int hg = ty & 3, dp = ty >> 2;
  const int pi_d = dp & 1, qd = dp >> 1;
  float4 acc[4][2];
#pragma unroll
  for (int r = 0; r < 4; ++r) acc[r][0] = acc[r][1] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int nkd = pi_d ? 2 : 1;
#pragma unroll 1
  for (int a = 0; a < nkd; ++a) {
    const int kd = pi_d ? (a == 0 ? 0 : 2) : 1;
    const int sd = pi_d ? (a == 0 ? qd + 1 : qd) : qd;
    const float4* base = tile + (sd * DB_SH + hg * 2) * DB_SW + tx;
    float4 xs[3][2];
#pragma unroll
    for (int rr = 0; rr < 3; ++rr) {
      xs[rr][0] = base[rr * DB_SW];
      xs[rr][1] = base[rr * DB_SW + 1];
    }
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const float4 wv = wsm[(kd * 3 + kh) * 3 + kw];
        const int e = (kw == 1) ? 0 : 1;
        const int sc = (kw == 0) ? 1 : 0;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int r = (kh == 1) ? 2 * half : 2 * half + 1;
          const int sr = (kh == 1) ? half : (kh == 0 ? half + 1 : half);
          fma4(acc[r][e], xs[sr][sc], wv);
        }
      }
  }
  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (A.bias) bias4 = ldg4(A.bias + c0);
  const int gd = d0 + dp;
  if (gd >= A.Db) return;
  float* yb = A.big + (long long)n * A.Db * A.Hb * A.Wb * A.ld_big + c0;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int gh = h0 + hg * 4 + r;
    if (gh >= A.Hb) break;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int gw = w0 + 2 * tx + e;
      if (gw >= A.Wb) continue;
      float* p = yb + (((long long)gd * A.Hb + gh) * A.Wb + gw) * A.ld_big;
      float4 v = make_float4(acc[r][e].x + bias4.x, acc[r][e].y + bias4.y, acc[r][e].z + bias4.z,
                             acc[r][e].w + bias4.w);
      if (A.accumulate) {
        const float4 o = *reinterpret_cast<const float4*>(p);
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
      }
      st4(p, v);
    }
  }
}

// =========================================================================================
// wgrad (persistent; blockIdx.y = channel chunk): dW[c][tap] += sum_o small[o,c]*big[S*o-1+tap,c]
// lanes as in wgrad3_s1_kernel: (kdkh, j) walks along w; 3 float4 accumulators per lane
// =========================================================================================
template <int S>
struct DwWgShape {
  static constexpr int TW = 32, TH = 6, TD = 2, NWARP = 4;
  static constexpr int PD = S * (TD - 1) + 3, PH = S * (TH - 1) + 3, PW = S * (TW - 1) + 3;
  static constexpr int PWP = PW | 1, PLANE_PAD = S == 1 ? 7 : 2;
  static constexpr int XPLANE = PH * PWP + PLANE_PAD;
  static constexpr int XTILE = PD * XPLANE + 8;
  static constexpr int YP = TW + 1, YTILE = TD * TH * YP + 8;
  static constexpr int THREADS = 32 * NWARP;
  static constexpr int NGROUPS = TD * TH / 3;
  static constexpr size_t SMEM = sizeof(float4) * (XTILE + YTILE) + sizeof(float) * (27 * 4 + 4);
};

template <int S>
__global__ void __launch_bounds__(DwWgShape<S>::THREADS)
    dw_wgrad3_kernel(const S2Args A, int ntiles) {
  using WS = DwWgShape<S>;
  constexpr int PW = WS::PW, PWP = WS::PWP, PH = WS::PH, TW = WS::TW, TH = WS::TH, TD = WS::TD;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* xt = reinterpret_cast<float4*>(smem_raw);
  float4* yt = xt + WS::XTILE;
  float* red = reinterpret_cast<float*>(yt + WS::YTILE);     // [27][4] + [4]
  const int c0 = blockIdx.y * 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool tap_lane = lane < 27;
  const int kdkh = tap_lane ? lane / 3 : 0;
  const int j = lane < 30 ? lane % 3 : 0;
  const bool active = lane < 30;
  const int kd = kdkh / 3, kh = kdkh % 3;
  float4 acc[3] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f),
                   make_float4(0.f, 0.f, 0.f, 0.f)};
  float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = threadIdx.x; i < 27 * 4 + 4; i += WS::THREADS) red[i] = 0.f;
#pragma unroll 1
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int b = tile;
    const int tw = b % A.tiles_w; b /= A.tiles_w;
    const int th = b % A.tiles_h; b /= A.tiles_h;
    const int td = b % A.tiles_d;
    const int n = b / A.tiles_d;
    const int w0 = tw * TW, h0 = th * TH, d0 = td * TD;
    __syncthreads();
    const float* xb = A.big + (long long)n * A.Db * A.Hb * A.Wb * A.ld_big + c0;
    // row-wise staging: one 64-bit row pointer per (d,h) row, 32-bit per-lane offsets
    for (int r = warp; r < WS::PD * PH; r += WS::NWARP) {
      const int pd = r / PH, ph = r - pd * PH;
      const int gd = S * d0 - 1 + pd, gh = S * h0 - 1 + ph;
      const bool rok = gd >= 0 && gd < A.Db && gh >= 0 && gh < A.Hb;
      const float* rp = xb + (long long)(gd * A.Hb + gh) * A.Wb * A.ld_big;
      float4* drow = xt + pd * WS::XPLANE + ph * PWP;
      for (int pw = lane; pw < PW; pw += 32) {
        const int gw = S * w0 - 1 + pw;
        const bool ok = rok && gw >= 0 && gw < A.Wb;
        dcp16(drow + pw, ok ? rp + gw * A.ld_big : A.big, ok);
      }
    }
    const float* yb = A.small + (long long)n * A.Ds * A.Hs * A.Ws * A.ld_small + c0;
    for (int i = threadIdx.x; i < TD * TH * TW; i += WS::THREADS) {
      const int pw = i % TW, r = i / TW;
      const int ph = r % TH, pd = r / TH;
      const int gd = d0 + pd, gh = h0 + ph, gw = w0 + pw;
      const bool ok = gd < A.Ds && gh < A.Hs && gw < A.Ws;
      const unsigned off = (unsigned)(((gd * A.Hs + gh) * A.Ws + gw) * A.ld_small);
      dcp16(&yt[r * WS::YP + pw], ok ? yb + off : A.small, ok);
    }
    dcp_wait_all();
    __syncthreads();
#pragma unroll 1
    for (int grp = warp; grp < WS::NGROUPS; grp += WS::NWARP) {
      if (!active) continue;
      const int row = grp * 3 + j;
      const int pd = row / TH, ph = row % TH;
      const float4* xr = xt + (S * pd + kd) * WS::XPLANE + (S * ph + kh) * PWP;
      const float4* yr = yt + (pd * TH + ph) * WS::YP;
#pragma unroll 8
      for (int w = 0; w < TW; ++w) {
        const float4 g = yr[w];
        fma4(acc[0], xr[S * w + 0], g);
        fma4(acc[1], xr[S * w + 1], g);
        fma4(acc[2], xr[S * w + 2], g);
        bsum.x += g.x; bsum.y += g.y; bsum.z += g.z; bsum.w += g.w;
      }
    }
  }
  if (tap_lane) {
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      float* r = red + (kdkh * 3 + kw) * 4;
      atomicAdd(r + 0, acc[kw].x); atomicAdd(r + 1, acc[kw].y);
      atomicAdd(r + 2, acc[kw].z); atomicAdd(r + 3, acc[kw].w);
    }
  } else if (active) {
    atomicAdd(&red[27 * 4 + 0], bsum.x); atomicAdd(&red[27 * 4 + 1], bsum.y);
    atomicAdd(&red[27 * 4 + 2], bsum.z); atomicAdd(&red[27 * 4 + 3], bsum.w);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 27 * 4; i += WS::THREADS)
    atomicAdd(A.dW + (long long)(c0 + i % 4) * 27 + i / 4, red[i]);
  if (A.dbias_small && threadIdx.x < 4) atomicAdd(A.dbias_small + c0 + threadIdx.x, red[27 * 4 + threadIdx.x]);
}

// =========================================================================================
// host side
// =========================================================================================
template <typename K>
static int set_smem(K kern, size_t bytes) {
  NAS3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return NAS3D_OK;
}

int tiled_dw_s1(bool flip, const TiledArgs& A0, int C, cudaStream_t st) {
  if (C % 4 || A0.ldx % 4 || A0.ldy % 4 || !aligned16(A0.x) || !aligned16(A0.y) || A0.W < 8)
    return NAS3D_ERR_UNSUPPORTED;
  if ((long long)A0.D * A0.H * A0.W * A0.ldx >= (1ll << 31)) return NAS3D_ERR_UNSUPPORTED;   // 32-bit staging offsets
  TiledArgs A = A0;
  A.tiles_w = (A.W + D1_TW - 1) / D1_TW;
  A.tiles_h = (A.H + D1_TH - 1) / D1_TH;
  A.tiles_d = (A.D + D1_TD - 1) / D1_TD;
  static bool done = false;
  if (!done) {
    int rc = set_smem(dw3_s1_kernel<false>, D1_SMEM); if (rc) return rc;
    rc = set_smem(dw3_s1_kernel<true>, D1_SMEM); if (rc) return rc;
    done = true;
  }
  const dim3 grid((unsigned)((long long)A.N * A.tiles_w * A.tiles_h * A.tiles_d), C / 4);
  if (flip) dw3_s1_kernel<true><<<grid, D1_THREADS, D1_SMEM, st>>>(A);
  else dw3_s1_kernel<false><<<grid, D1_THREADS, D1_SMEM, st>>>(A);
  return launched("dw3_s1");
}

static bool dw_s2_ok(const S2Args& A) {
  if ((long long)A.Db * A.Hb * A.Wb * A.ld_big >= (1ll << 31)) return false;   // 32-bit staging offsets
  return A.Cb == A.Cs && A.Cb % 4 == 0 && A.ld_big % 4 == 0 && A.ld_small % 4 == 0 &&
         aligned16(A.big) && aligned16(A.small) && A.Db == 2 * A.Ds && A.Hb == 2 * A.Hs &&
         A.Wb == 2 * A.Ws && A.Ws >= 8;
}

int tiled_dw_s2_sfb(const S2Args& A0, cudaStream_t st) {
  if (!dw_s2_ok(A0)) return NAS3D_ERR_UNSUPPORTED;
  S2Args A = A0;
  A.tiles_w = (A.Ws + DS_TW - 1) / DS_TW;
  A.tiles_h = (A.Hs + DS_TH - 1) / DS_TH;
  A.tiles_d = (A.Ds + DS_TD - 1) / DS_TD;
  static bool done = false;
  if (!done) { int rc = set_smem(dw3_s2_sfb_kernel, DS_SMEM); if (rc) return rc; done = true; }
  const dim3 grid((unsigned)((long long)A.N * A.tiles_w * A.tiles_h * A.tiles_d), A.Cb / 4);
  dw3_s2_sfb_kernel<<<grid, DS_THREADS, DS_SMEM, st>>>(A);
  return launched("dw3_s2_sfb");
}

int tiled_dw_s2_bfs(const S2Args& A0, cudaStream_t st) {
  if (!dw_s2_ok(A0)) return NAS3D_ERR_UNSUPPORTED;
  S2Args A = A0;
  A.tiles_w = (A.Wb + DB_TWB - 1) / DB_TWB;
  A.tiles_h = (A.Hb + DB_THB - 1) / DB_THB;
  A.tiles_d = (A.Db + DB_TDB - 1) / DB_TDB;
  static bool done = false;
  if (!done) { int rc = set_smem(dw3_s2_bfs_kernel, DB_SMEM); if (rc) return rc; done = true; }
  const dim3 grid((unsigned)((long long)A.N * A.tiles_w * A.tiles_h * A.tiles_d), A.Cb / 4);
  dw3_s2_bfs_kernel<<<grid, DB_THREADS, DB_SMEM, st>>>(A);
  return launched("dw3_s2_bfs");
}

template <int S>
static int launch_dw_wgrad(S2Args A, cudaStream_t st) {
  using WS = DwWgShape<S>;
  A.tiles_w = (A.Ws + WS::TW - 1) / WS::TW;
  A.tiles_h = (A.Hs + WS::TH - 1) / WS::TH;
  A.tiles_d = (A.Ds + WS::TD - 1) / WS::TD;
  static int occ = 0;
  if (!occ) {
    int rc = set_smem(dw_wgrad3_kernel<S>, WS::SMEM); if (rc) return rc;
    NAS3D_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dw_wgrad3_kernel<S>, WS::THREADS, WS::SMEM));
    if (occ < 1) occ = 1;
  }
  const long long ntiles = (long long)A.N * A.tiles_w * A.tiles_h * A.tiles_d;
  const int chunks = A.Cb / 4;
  long long gx = (long long)kNumSMs * occ / chunks;
  if (gx < 1) gx = 1;
  if (gx > ntiles) gx = ntiles;
  dw_wgrad3_kernel<S><<<dim3((unsigned)gx, chunks), WS::THREADS, WS::SMEM, st>>>(A, (int)ntiles);
  return launched("dw_wgrad3");
}

// stride 1: big and small have the same extents
int tiled_dw_wgrad(const S2Args& A, int stride, cudaStream_t st) {
  if (A.Cb != A.Cs || A.Cb % 4 || A.ld_big % 4 || A.ld_small % 4 || !aligned16(A.big) ||
      !aligned16(A.small) || A.Ws < 8)
    return NAS3D_ERR_UNSUPPORTED;
  if ((long long)A.Db * A.Hb * A.Wb * A.ld_big >= (1ll << 31)) return NAS3D_ERR_UNSUPPORTED;
  if (stride == 1 && A.Db == A.Ds && A.Hb == A.Hs && A.Wb == A.Ws) return launch_dw_wgrad<1>(A, st);
  if (stride == 2 && dw_s2_ok(A)) {
    const int rc = tiled_dw_s2_wgrad_tma(A, st);
    if (rc != NAS3D_ERR_UNSUPPORTED) return rc;
    return launch_dw_wgrad<2>(A, st);
  }
  return NAS3D_ERR_UNSUPPORTED;
}

}  // namespace nas3d
