// Tiled 3x3x3 STRIDE-2 (dilation 1, pad 1) convolutions for the narrow layers (C = 4, 8 and the
// 4->12 stem): the down_conv / up_conv family and stem1.  Conv view (nas3d_b200.h):
//     big_pos = 2*small_pos - 1 + tap        W[Cs][Cb][27]
//
//   conv3_s2_sfb : small[o]  = bias + sum_tap big[2o-1+tap] W[.][.][tap]   (Conv3d fwd, ConvT dgrad)
//   conv3_s2_bfs : big[i]  (+)= bias + sum over taps with i+1-tap even of small[(i+1-tap)/2] W
//                                                                           (ConvT fwd, Conv3d dgrad)
//   wgrad3_s2    : dW[tap]  += sum_o small[o] (x) big[2o-1+tap]
//
// Same machinery as conv_tiled.cu (cp.async staged tiles with zero-filled halo, 4 output
// channels x several voxels per thread, weights broadcast from smem, persistent register
// accumulation for wgrad) with the stride folded into the smem layout:
//  * sfb: the big tile row is stored de-interleaved (even columns | odd columns) so that the 32
//    lanes (consecutive small w) still read consecutive float4 for every kw;
//  * bfs: a thread produces the 2 big voxels (2w, 2w+1) of 4 big rows from a 3x2 patch of small
//    voxels; per parity class only the taps that land on the lattice are visited (27/8 per voxel);
//  * wgrad: a lane walks along small w with a 3-wide window over the big row (2 new loads/step).
// HBM/L2-bound by construction (8 big voxels per small voxel): AI = 6.75 FLOP/B at C = 4.
#include "common.cuh"
#include "conv_tiled.h"

namespace nas3d {

__device__ __forceinline__ void cp16(void* smem_dst, const void* gsrc, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// =========================================================================================
// small-from-big
// =========================================================================================
template <int CIN, int COUT, int HG, int DG>
struct S2FwdShape {
  static constexpr int C4I = CIN / 4, C4O = COUT / 4;
  static constexpr int TW = 32, TH = 4 * HG, TD = DG;
  static constexpr int PD = 2 * TD + 1, PH = 2 * TH + 1, PWS = 2 * TW + 1, EVEN_N = TW + 1;
  static constexpr int PLANE = PD * PH * PWS;
  static constexpr int THREADS = 32 * HG * DG * C4O;
  static constexpr size_t SMEM = sizeof(float4) * PLANE * C4I + sizeof(float) * 27 * CIN * COUT;
};

template <int CIN, int COUT, int HG, int DG>
__global__ void __launch_bounds__(S2FwdShape<CIN, COUT, HG, DG>::THREADS)
    conv3_s2_sfb_kernel(const S2Args A) {
  using TS = S2FwdShape<CIN, COUT, HG, DG>;
  constexpr int C4I = TS::C4I, C4O = TS::C4O, PH = TS::PH, PWS = TS::PWS, EVEN_N = TS::EVEN_N;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* tile = reinterpret_cast<float4*>(smem_raw);
  float* wsm = reinterpret_cast<float*>(smem_raw + sizeof(float4) * TS::PLANE * C4I);
  __shared__ double sm_mom[2 * COUT];
  if (threadIdx.x < 2 * COUT) sm_mom[threadIdx.x] = 0.0;

  int b = blockIdx.x;
  const int tw = b % A.tiles_w; b /= A.tiles_w;
  const int th = b % A.tiles_h; b /= A.tiles_h;
  const int td = b % A.tiles_d;
  const int n = b / A.tiles_d;
  const int w0 = tw * TS::TW, h0 = th * TS::TH, d0 = td * TS::TD;     // small coordinates

  // wsm[tap][ci][co] <- W[co(cs)][ci(cb)][tap]
  for (int i = threadIdx.x; i < 27 * CIN * COUT; i += TS::THREADS) {
    const int co = i % COUT, ci = (i / COUT) % CIN, tap = i / (COUT * CIN);
    wsm[i] = __ldg(A.w + ((long long)co * CIN + ci) * 27 + tap);
  }
  const float* xb = A.big + (long long)n * A.Db * A.Hb * A.Wb * A.ld_big;
  for (int i = threadIdx.x; i < TS::PLANE * C4I; i += TS::THREADS) {
    const int cc = i % C4I;
    int r = i / C4I;
    const int pwi = r % PWS; r /= PWS;
    const int ph = r % PH;
    const int pd = r / PH;
    const int bw = pwi < EVEN_N ? 2 * pwi : 2 * (pwi - EVEN_N) + 1;
    const int gd = 2 * d0 - 1 + pd, gh = 2 * h0 - 1 + ph, gw = 2 * w0 - 1 + bw;
    const bool ok = gd >= 0 && gd < A.Db && gh >= 0 && gh < A.Hb && gw >= 0 && gw < A.Wb;
    // 32-bit in-sample offset (host guarantees samples < 2^31 floats)
    const unsigned off = (unsigned)(((gd * A.Hb + gh) * A.Wb + gw) * A.ld_big + cc * 4);
    cp16(&tile[cc * TS::PLANE + (pd * PH + ph) * PWS + pwi], ok ? xb + off : A.big, ok);
  }
  cp_wait_all();
  __syncthreads();

  const int tx = threadIdx.x & 31;
  int ty = threadIdx.x >> 5;
  const int cg = ty % C4O; ty /= C4O;
  const int hg = ty % HG;
  const int dg = ty / HG;

  float2 acc[4][2];      // [h][channel pair]: FFMA2 accumulators (common.cuh)
#pragma unroll
  for (int h = 0; h < 4; ++h)
#pragma unroll
    for (int c = 0; c < 2; ++c) acc[h][c] = make_float2(0.f, 0.f);

#pragma unroll 1
  for (int kd = 0; kd < 3; ++kd) {
#pragma unroll 1
    for (int kw = 0; kw < 3; ++kw) {
      const int woff = (kw == 1) ? EVEN_N + tx : tx + (kw >> 1);
#pragma unroll 1
      for (int cc = 0; cc < C4I; ++cc) {
        float4 wr[3][4];
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int ci = 0; ci < 4; ++ci)
            wr[kh][ci] = *reinterpret_cast<const float4*>(
                wsm + (((kd * 3 + kh) * 3 + kw) * CIN + cc * 4 + ci) * COUT + cg * 4);
        const float4* col = tile + cc * TS::PLANE + ((2 * dg + kd) * PH + 2 * (hg * 4)) * PWS + woff;
        float4 xin[9];
#pragma unroll
        for (int r = 0; r < 9; ++r) xin[r] = col[r * PWS];
#pragma unroll
        for (int oh = 0; oh < 4; ++oh)
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) dot4x4(acc[oh], xin[2 * oh + kh], wr[kh]);
      }
    }
  }

  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (A.bias) bias4 = ldg4(A.bias + cg * 4);
  const int gw = w0 + tx, gd = d0 + dg;
  float ms[4] = {0.f, 0.f, 0.f, 0.f}, mq[4] = {0.f, 0.f, 0.f, 0.f};
  if (gw < A.Ws && gd < A.Ds) {
    float* yb = A.small + (long long)n * A.Ds * A.Hs * A.Ws * A.ld_small + cg * 4;
#pragma unroll
    for (int oh = 0; oh < 4; ++oh) {
      const int gh = h0 + hg * 4 + oh;
      if (gh >= A.Hs) break;
      float* p = yb + (((long long)gd * A.Hs + gh) * A.Ws + gw) * A.ld_small;
      float4 v = make_float4(acc[oh][0].x + bias4.x, acc[oh][0].y + bias4.y, acc[oh][1].x + bias4.z,
                             acc[oh][1].y + bias4.w);
      if (A.accumulate) {
        const float4 o = *reinterpret_cast<const float4*>(p);
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
      }
      st4(p, v);
      ms[0] += v.x; ms[1] += v.y; ms[2] += v.z; ms[3] += v.w;
      mq[0] += v.x * v.x; mq[1] += v.y * v.y; mq[2] += v.z * v.z; mq[3] += v.w * v.w;
    }
  }
  if (A.moments) {
    warp_moments_add(sm_mom, cg * 4, ms, mq);
    __syncthreads();
    cta_moments_flush(sm_mom, A.moments, n, COUT, TS::THREADS);
  }
}

// =========================================================================================
// big-from-small (C = Cs = Cb)
// =========================================================================================
template <int C>
struct S2BfsShape {
  static constexpr int C4 = C / 4;
  static constexpr int TDB = 4, THB = 16, TWB = 64;               // big tile
  static constexpr int SD = TDB / 2 + 1, SH = THB / 2 + 1, SW = TWB / 2 + 1;   // small tile
  static constexpr int PLANE = SD * SH * SW;
  static constexpr int THREADS = 32 * (THB / 4) * TDB * C4;
  static constexpr size_t SMEM = sizeof(float4) * PLANE * C4 + sizeof(float) * 27 * C * C;
};

template <int C>
__global__ void __launch_bounds__(S2BfsShape<C>::THREADS) conv3_s2_bfs_kernel(const S2Args A) {
  using TS = S2BfsShape<C>;
  constexpr int C4 = TS::C4, SH = TS::SH, SW = TS::SW;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* tile = reinterpret_cast<float4*>(smem_raw);
  float* wsm = reinterpret_cast<float*>(smem_raw + sizeof(float4) * TS::PLANE * C4);
  __shared__ double sm_mom[2 * C];
  if (threadIdx.x < 2 * C) sm_mom[threadIdx.x] = 0.0;

  int b = blockIdx.x;
  const int tw = b % A.tiles_w; b /= A.tiles_w;
  const int th = b % A.tiles_h; b /= A.tiles_h;
  const int td = b % A.tiles_d;
  const int n = b / A.tiles_d;
  const int w0 = tw * TS::TWB, h0 = th * TS::THB, d0 = td * TS::TDB;   // big coordinates (even)

  // wsm[tap][cs(red)][cb(prod)] <- W[cs][cb][tap]
  for (int i = threadIdx.x; i < 27 * C * C; i += TS::THREADS) {
    const int cb = i % C, cs = (i / C) % C, tap = i / (C * C);
    wsm[i] = __ldg(A.w + ((long long)cs * C + cb) * 27 + tap);
  }
  const float* sb = A.small + (long long)n * A.Ds * A.Hs * A.Ws * A.ld_small;
  for (int i = threadIdx.x; i < TS::PLANE * C4; i += TS::THREADS) {
    const int cc = i % C4;
    int r = i / C4;
    const int pw = r % SW; r /= SW;
    const int ph = r % SH;
    const int pd = r / SH;
    const int gd = d0 / 2 + pd, gh = h0 / 2 + ph, gw = w0 / 2 + pw;
    const bool ok = gd < A.Ds && gh < A.Hs && gw < A.Ws;
    const unsigned off = (unsigned)(((gd * A.Hs + gh) * A.Ws + gw) * A.ld_small + cc * 4);
    cp16(&tile[cc * TS::PLANE + (pd * SH + ph) * SW + pw], ok ? sb + off : A.small, ok);
  }
  cp_wait_all();
  __syncthreads();

  const int tx = threadIdx.x & 31;
  int ty = threadIdx.x >> 5;
  const int cg = ty % C4; ty /= C4;
  const int hg = ty % (TS::THB / 4);
  const int dp = ty / (TS::THB / 4);            // big plane inside the tile
  const int pi_d = dp & 1, qd = dp >> 1;

  float2 acc[4][2][2];   // [big row r][big col e][co pair]: FFMA2 accumulators
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int e = 0; e < 2; ++e)
#pragma unroll
      for (int c = 0; c < 2; ++c) acc[r][e][c] = make_float2(0.f, 0.f);

  // taps along d for this plane: even plane -> kd = 1 @ qd ; odd plane -> kd = 0 @ qd+1, kd = 2 @ qd
  const int nkd = pi_d ? 2 : 1;
#pragma unroll 1
  for (int a = 0; a < nkd; ++a) {
    const int kd = pi_d ? (a == 0 ? 0 : 2) : 1;
    const int sd = pi_d ? (a == 0 ? qd + 1 : qd) : qd;
#pragma unroll 1
    for (int cc = 0; cc < C4; ++cc) {
      // 3 small rows x 2 small cols of my patch
      const float4* base = tile + cc * TS::PLANE + (sd * SH + hg * 2) * SW + tx;
      float4 xs[3][2];
#pragma unroll
      for (int rr = 0; rr < 3; ++rr) {
        xs[rr][0] = base[rr * SW];
        xs[rr][1] = base[rr * SW + 1];
      }
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          float4 wv[4];
#pragma unroll
          for (int ci = 0; ci < 4; ++ci)
            wv[ci] = *reinterpret_cast<const float4*>(
                wsm + (((kd * 3 + kh) * 3 + kw) * C + cc * 4 + ci) * C + cg * 4);
          // big rows touched by kh: kh==1 -> even rows (r = 0,2) from small row r/2
          //                         kh==0 -> odd rows (r = 1,3) from small row (r+1)/2
          //                         kh==2 -> odd rows (r = 1,3) from small row (r-1)/2
          const int e = (kw == 1) ? 0 : 1;            // big col parity
          const int sc = (kw == 0) ? 1 : 0;           // small col offset
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int r = (kh == 1) ? 2 * half : 2 * half + 1;
            const int sr = (kh == 1) ? half : (kh == 0 ? half + 1 : half);
            dot4x4(acc[r][e], xs[sr][sc], wv);
          }
        }
      }
    }
  }

  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (A.bias) bias4 = ldg4(A.bias + cg * 4);
  const int gd = d0 + dp;
  float ms[4] = {0.f, 0.f, 0.f, 0.f}, mq[4] = {0.f, 0.f, 0.f, 0.f};
  if (gd < A.Db) {
    float* yb = A.big + (long long)n * A.Db * A.Hb * A.Wb * A.ld_big + cg * 4;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int gh = h0 + hg * 4 + r;
      if (gh >= A.Hb) break;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int gw = w0 + 2 * tx + e;
        if (gw >= A.Wb) continue;
        float* p = yb + (((long long)gd * A.Hb + gh) * A.Wb + gw) * A.ld_big;
        float4 v = make_float4(acc[r][e][0].x + bias4.x, acc[r][e][0].y + bias4.y,
                               acc[r][e][1].x + bias4.z, acc[r][e][1].y + bias4.w);
        if (A.accumulate) {
          const float4 o = *reinterpret_cast<const float4*>(p);
          v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
        }
        st4(p, v);
        ms[0] += v.x; ms[1] += v.y; ms[2] += v.z; ms[3] += v.w;
        mq[0] += v.x * v.x; mq[1] += v.y * v.y; mq[2] += v.z * v.z; mq[3] += v.w * v.w;
      }
    }
  }
  if (A.moments) {
    warp_moments_add(sm_mom, cg * 4, ms, mq);
    __syncthreads();
    cta_moments_flush(sm_mom, A.moments, n, C, TS::THREADS);
  }
}

// =========================================================================================
// wgrad (persistent, one (cb-chunk, cs-chunk) pair per blockIdx.y)
// =========================================================================================
template <int TWT>
struct WgS2Shape {
  static constexpr int TW = TWT, TH = 6, TD = 2, NWARP = 4;     // TW 32 | 16 | 8 (narrow volumes)
  static constexpr int PD = 2 * TD + 1, PH = 2 * TH + 1, PW = 2 * TW + 1;
  static constexpr int PWP = PW, PLANE_PAD = 2;        // bank-conflict search (see conv_tiled.cu)
  static constexpr int XPLANE = PH * PWP + PLANE_PAD;
  static constexpr int XTILE = PD * XPLANE + 8;
  static constexpr int YP = TW + 1;
  static constexpr int YTILE = TD * TH * YP + 8;
  static constexpr int THREADS = 32 * NWARP;
  static constexpr int ROWS = TD * TH, NGROUPS = ROWS / 3;
  static_assert(NGROUPS % NWARP == 0, "rows per warp");
  static constexpr size_t SMEM = sizeof(float4) * (XTILE + YTILE) + sizeof(float) * (27 * 16 + 4);
};

// Row-wise staging / BIAS template: see wgrad3_s1_kernel (conv_tiled.cu).
template <int TWT, bool BIAS>
__global__ void __launch_bounds__(WgS2Shape<TWT>::THREADS)
    wgrad3_s2_kernel(const S2Args A, int ntiles) {
  using WS = WgS2Shape<TWT>;
  constexpr int PW = WS::PW, PWP = WS::PWP, PH = WS::PH, TW = WS::TW, TH = WS::TH, TD = WS::TD;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* xt = reinterpret_cast<float4*>(smem_raw);
  float4* yt = xt + WS::XTILE;
  float* red = reinterpret_cast<float*>(yt + WS::YTILE);

  const int C4B = A.Cb / 4, C4S = A.Cs / 4;
  const int cic = blockIdx.y / C4S, coc = blockIdx.y % C4S;    // cic: big chunk, coc: small chunk
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool tap_lane = lane < 27;
  const int kdkh = tap_lane ? lane / 3 : 0;
  const int j = lane < 30 ? lane % 3 : 0;
  const bool active = lane < 30;
  const int kd = kdkh / 3, kh = kdkh % 3;
  (void)C4B;

  float2 acc[3][4][2];   // [kw][cb][cs pair]: FFMA2 accumulators
#pragma unroll
  for (int t = 0; t < 3; ++t)
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 2; ++c) acc[t][a][c] = make_float2(0.f, 0.f);
  float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = threadIdx.x; i < 27 * 16 + 4; i += WS::THREADS) red[i] = 0.f;

#pragma unroll 1
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int b = tile;
    const int tw = b % A.tiles_w; b /= A.tiles_w;
    const int th = b % A.tiles_h; b /= A.tiles_h;
    const int td = b % A.tiles_d;
    const int n = b / A.tiles_d;
    const int w0 = tw * TW, h0 = th * TH, d0 = td * TD;     // small coordinates
    __syncthreads();
    const float* xb = A.big + (long long)n * A.Db * A.Hb * A.Wb * A.ld_big + cic * 4;
    for (int r = warp; r < WS::PD * PH; r += WS::NWARP) {
      const int pd = r / PH, ph = r - pd * PH;
      const int gd = 2 * d0 - 1 + pd, gh = 2 * h0 - 1 + ph;
      const bool rok = gd >= 0 && gd < A.Db && gh >= 0 && gh < A.Hb;
      const float* rp = xb + (long long)(gd * A.Hb + gh) * A.Wb * A.ld_big;
      float4* drow = xt + pd * WS::XPLANE + ph * PWP;
      for (int pw = lane; pw < PW; pw += 32) {
        const int gw = 2 * w0 - 1 + pw;
        const bool ok = rok && gw >= 0 && gw < A.Wb;
        cp16(drow + pw, ok ? rp + gw * A.ld_big : A.big, ok);
      }
    }
    const float* yb = A.small + (long long)n * A.Ds * A.Hs * A.Ws * A.ld_small + coc * 4;
    for (int i = threadIdx.x; i < TD * TH * TW; i += WS::THREADS) {
      const int pw = i % TW, r = i / TW;
      const int ph = r % TH, pd = r / TH;
      const int gd = d0 + pd, gh = h0 + ph, gw = w0 + pw;
      const bool ok = gd < A.Ds && gh < A.Hs && gw < A.Ws;
      const unsigned off = (unsigned)(((gd * A.Hs + gh) * A.Ws + gw) * A.ld_small);
      cp16(&yt[r * WS::YP + pw], ok ? yb + off : A.small, ok);
    }
    cp_wait_all();
    __syncthreads();

#pragma unroll 1
    for (int grp = warp; grp < WS::NGROUPS; grp += WS::NWARP) {
      if (!active) continue;
      const int row = grp * 3 + j;
      const int pd = row / TH, ph = row % TH;
      const float4* xr = xt + (2 * pd + kd) * WS::XPLANE + (2 * ph + kh) * PWP;
      const float4* yr = yt + (pd * TH + ph) * WS::YP;
      float4 xa = xr[0];
#pragma unroll 4
      for (int w = 0; w < TW; ++w) {
        const float4 xb2 = xr[2 * w + 1];
        const float4 xc = xr[2 * w + 2];
        const float4 g = yr[w];
        const float4 xv3[3] = {xa, xb2, xc};
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) outer4(acc[kw], xv3[kw], g);
        if (BIAS) { bsum.x += g.x; bsum.y += g.y; bsum.z += g.z; bsum.w += g.w; }
        xa = xc;
      }
    }
  }

  if (tap_lane) {
#pragma unroll
    for (int kw = 0; kw < 3; ++kw)
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c)
          atomicAdd(&red[((kdkh * 3 + kw) * 4 + a) * 4 + c], (c & 1) ? acc[kw][a][c >> 1].y : acc[kw][a][c >> 1].x);
  } else if (BIAS && active) {
    atomicAdd(&red[27 * 16 + 0], bsum.x); atomicAdd(&red[27 * 16 + 1], bsum.y);
    atomicAdd(&red[27 * 16 + 2], bsum.z); atomicAdd(&red[27 * 16 + 3], bsum.w);
  }
  __syncthreads();
  // red[tap][cb][cs] -> dW[cs][cb][tap]
  for (int i = threadIdx.x; i < 27 * 16; i += WS::THREADS) {
    const int cs = coc * 4 + i % 4, cb = cic * 4 + (i / 4) % 4, t = i / 16;
    atomicAdd(A.dW + ((long long)cs * A.Cb + cb) * 27 + t, red[i]);
  }
  if (BIAS && A.dbias_small && cic == 0 && threadIdx.x < 4)
    atomicAdd(A.dbias_small + coc * 4 + threadIdx.x, red[27 * 16 + threadIdx.x]);
}

// =========================================================================================
// host side
// =========================================================================================
template <int CIN, int COUT, int HG, int DG>
static int launch_sfb(S2Args A, cudaStream_t st) {
  using TS = S2FwdShape<CIN, COUT, HG, DG>;
  A.tiles_w = (A.Ws + TS::TW - 1) / TS::TW;
  A.tiles_h = (A.Hs + TS::TH - 1) / TS::TH;
  A.tiles_d = (A.Ds + TS::TD - 1) / TS::TD;
  auto kern = conv3_s2_sfb_kernel<CIN, COUT, HG, DG>;
  static bool attr_done = false;
  if (!attr_done) {
    NAS3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TS::SMEM));
    attr_done = true;
  }
  const long long blocks = (long long)A.N * A.tiles_w * A.tiles_h * A.tiles_d;
  kern<<<(unsigned)blocks, TS::THREADS, TS::SMEM, st>>>(A);
  return launched("conv3_s2_sfb");
}

template <int C>
static int launch_bfs(S2Args A, cudaStream_t st) {
  using TS = S2BfsShape<C>;
  A.tiles_w = (A.Wb + TS::TWB - 1) / TS::TWB;
  A.tiles_h = (A.Hb + TS::THB - 1) / TS::THB;
  A.tiles_d = (A.Db + TS::TDB - 1) / TS::TDB;
  auto kern = conv3_s2_bfs_kernel<C>;
  static bool attr_done = false;
  if (!attr_done) {
    NAS3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TS::SMEM));
    attr_done = true;
  }
  const long long blocks = (long long)A.N * A.tiles_w * A.tiles_h * A.tiles_d;
  kern<<<(unsigned)blocks, TS::THREADS, TS::SMEM, st>>>(A);
  return launched("conv3_s2_bfs");
}

static bool s2_common_ok(const S2Args& A) {
  if ((long long)A.Db * A.Hb * A.Wb * A.ld_big >= (1ll << 31)) return false;   // 32-bit staging offsets
  return A.ld_big % 4 == 0 && A.ld_small % 4 == 0 && aligned16(A.big) && aligned16(A.small) &&
         A.Db == 2 * A.Ds && A.Hb == 2 * A.Hs && A.Wb == 2 * A.Ws && A.Ws >= 8;
}

int tiled_s2_sfb(const S2Args& A, cudaStream_t st) {
  if (!s2_common_ok(A)) return NAS3D_ERR_UNSUPPORTED;
  if (A.Cb == 4 && A.Cs == 12) return launch_sfb<4, 12, 2, 2>(A, st);
  if (A.Cb == 4 && A.Cs == 4) return launch_sfb<4, 4, 2, 2>(A, st);
  if (A.Cb == 8 && A.Cs == 8) return launch_sfb<8, 8, 2, 2>(A, st);
  return NAS3D_ERR_UNSUPPORTED;
}

int tiled_s2_bfs(const S2Args& A, cudaStream_t st) {
  if (!s2_common_ok(A)) return NAS3D_ERR_UNSUPPORTED;
  if (A.Cb == 4 && A.Cs == 4) return launch_bfs<4>(A, st);
  if (A.Cb == 8 && A.Cs == 8) return launch_bfs<8>(A, st);
  return NAS3D_ERR_UNSUPPORTED;
}

template <int TWT, bool BIAS>
static int launch_s2_wgrad_b(S2Args A, cudaStream_t st) {
  using WS = WgS2Shape<TWT>;
  A.tiles_w = (A.Ws + WS::TW - 1) / WS::TW;
  A.tiles_h = (A.Hs + WS::TH - 1) / WS::TH;
  A.tiles_d = (A.Ds + WS::TD - 1) / WS::TD;
  static int occ = 0;
  if (!occ) {
    NAS3D_CUDA(cudaFuncSetAttribute(wgrad3_s2_kernel<TWT, BIAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS::SMEM));
    NAS3D_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, wgrad3_s2_kernel<TWT, BIAS>, WS::THREADS, WS::SMEM));
    if (occ < 1) occ = 1;
  }
  const long long ntiles = (long long)A.N * A.tiles_w * A.tiles_h * A.tiles_d;
  const int pairs = (A.Cb / 4) * (A.Cs / 4);
  long long gx = (long long)kNumSMs * occ / pairs;
  if (gx < 1) gx = 1;
  if (gx > ntiles) gx = ntiles;
  wgrad3_s2_kernel<TWT, BIAS><<<dim3((unsigned)gx, pairs), WS::THREADS, WS::SMEM, st>>>(A, (int)ntiles);
  return launched("wgrad3_s2");
}

template <int TWT>
static int launch_s2_wgrad(const S2Args& A, cudaStream_t st) {
  return A.dbias_small ? launch_s2_wgrad_b<TWT, true>(A, st) : launch_s2_wgrad_b<TWT, false>(A, st);
}

int tiled_s2_wgrad(const S2Args& A, cudaStream_t st) {
  const bool ok = A.ld_big % 4 == 0 && A.ld_small % 4 == 0 && aligned16(A.big) && aligned16(A.small) &&
                  A.Db == 2 * A.Ds && A.Hb == 2 * A.Hs && A.Wb == 2 * A.Ws && A.Ws >= 2;
  if (!ok || A.Cb % 4 || A.Cs % 4 || A.Cb > 64 || A.Cs > 64) return NAS3D_ERR_UNSUPPORTED;
  if ((long long)A.Db * A.Hb * A.Wb * A.ld_big >= (1ll << 31)) return NAS3D_ERR_UNSUPPORTED;
  if (A.Ws <= 8) return launch_s2_wgrad<8>(A, st);
  if (A.Ws <= 16) return launch_s2_wgrad<16>(A, st);
  return launch_s2_wgrad<32>(A, st);
}

}  // namespace nas3d
