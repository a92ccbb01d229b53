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
#include <cuda.h>
#include <string.h>
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
  // register cap for two resident CTAs of the 384-thread 4->12 stem kernel: with one CTA per SM
  // its staging and FMA phases never overlapped (tools/s2_micro.py: 289 -> 205 us); 0 = no cap
  static constexpr int MINB = THREADS > 256 ? 2 : 0;
  static constexpr size_t SMEM = sizeof(float4) * PLANE * C4I + sizeof(float) * 27 * CIN * COUT;
};

template <int CIN, int COUT, int HG, int DG>
__global__ void __launch_bounds__(S2FwdShape<CIN, COUT, HG, DG>::THREADS, S2FwdShape<CIN, COUT, HG, DG>::MINB)
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
  // big tile; C = 8: 8 rows (512 threads, two CTAs per SM) - 58 -> 46 us at 8 x 64^3 -> 32^3
  static constexpr int TDB = 4, THB = C == 8 ? 8 : 16, TWB = 64;
  static constexpr int SD = TDB / 2 + 1, SH = THB / 2 + 1, SW = TWB / 2 + 1;   // small tile
  static constexpr int PLANE = SD * SH * SW;
  static constexpr int THREADS = 32 * (THB / 4) * TDB * C4;
  // 512 threads at 80-84 registers left one CTA per SM; capped at 64 two are resident
  // (tools/conv_micro.py stride 2, 8 x 128^3: dgrad 143.6 -> 110.7 us, profiles/r4j_minb_micro.txt)
  static constexpr int MINB = 2;
  static constexpr size_t SMEM = sizeof(float4) * PLANE * C4 + sizeof(float) * 27 * C * C;
};

template <int C>
__global__ void __launch_bounds__(S2BfsShape<C>::THREADS, S2BfsShape<C>::MINB) conv3_s2_bfs_kernel(const S2Args A) {
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
// wgrad, TMA-staged and double-buffered (the default; the cp.async kernel above is the fallback
// when no tensor map can be made)
// =========================================================================================
// ncu of the cp.async kernel (profiles/r3x_ncu_full_s2_family.csv): FFMA2 is 23 % of the issued
// instructions, the per-element LDGSTS staging (47 instructions and 10.6 shared-memory wavefronts
// per copy) costs as many shared-memory cycles as all operand reads of the FMA loop, nothing
// overlaps (stage -> wait -> barrier -> FMAs), and the x reads take 6 wavefronts instead of 4.
// Here one thread issues two cp.async.bulk.tensor copies per tile (haloed big tile, dy tile) into
// the buffer the CTA is NOT computing on; out-of-volume voxels are zero-filled by the TMA unit
// (= the conv padding, and dy = 0 outside contributes nothing), so the instruction stream is
// the FMA loop alone.  CSV = 2 keeps 8 output channels per lane (96 accumulators): twice the
// FMAs per x operand read, and the x tile is re-staged half as often.
//
// lane -> (kd, kh, row j): with the dense TMA tile (row pitch 2TW+1 = 1, plane pitch
// 13(2TW+1) = 5 mod 8 sixteen-byte bank groups) this permutation gives every quarter-warp 8
// distinct bank groups (or the same address); bias / idle lanes duplicate a neighbour's address.
// byte = kd | kh << 2 | j << 4 | role << 6   (role 0: tap lane, 1: bias lane, 2: idle)
#define S2T(kd, kh, j) ((kd) | (kh) << 2 | (j) << 4)
__constant__ unsigned char kS2LaneTap[32] = {
    S2T(2, 1, 2), S2T(2, 1, 1), S2T(2, 0, 1), S2T(0, 0, 0), S2T(1, 1, 0), S2T(0, 1, 0), S2T(2, 2, 0), S2T(0, 1, 1),
    S2T(2, 2, 1), S2T(0, 0, 1), S2T(1, 1, 1), S2T(0, 2, 1), S2T(0, 2, 0), S2T(2, 0, 2), S2T(0, 0, 2), S2T(2, 0, 2) | 64,
    S2T(1, 0, 2), S2T(0, 1, 2), S2T(1, 2, 2), S2T(1, 2, 1), S2T(1, 1, 2), S2T(0, 2, 2), S2T(1, 2, 1) | 64, 128,
    S2T(1, 0, 1), S2T(2, 2, 2), S2T(1, 0, 0), S2T(2, 1, 0), S2T(1, 2, 0), S2T(2, 0, 0), S2T(1, 0, 0) | 64, 128};
#undef S2T

template <int TWT, int CSV>
struct WgS2TmaShape {
  // Tile ring: as many stages as let two CTAs share an SM.  Measured (profiles/r4_s2_wgrad.txt):
  // the 4-channel kernel is bound by shared-memory bandwidth (11 LDS wavefronts per 24 FFMA2 +
  // the TMA fill: ncu r4c 56 % + ~25 % of the pipe), so neither a third resident CTA (2 stages)
  // nor 12 output channels per lane (207 registers) beat this shape.
  static constexpr int NST = CSV == 1 ? 3 : 2;
  static constexpr int TW = TWT, TH = 6, TD = 2, NWARP = 4;    // consumer warps; + 1 producer warp
  static constexpr int THREADS = 32 * (NWARP + 1);
  static constexpr int PD = 2 * TD + 1, PH = 2 * TH + 1, PW = 2 * TW + 1;
  static constexpr int XPLANE = PH * PW;
  static constexpr int XBYTES = PD * XPLANE * 16;
  static constexpr int YW = TW + 1;               // spare column: the warp's 3 dy rows start in different bank groups
  static constexpr int YROW = YW * CSV;           // float4 per dy row
  static constexpr int YBYTES = TD * TH * YROW * 16;
  static constexpr int XPAD = (XBYTES + 127) / 128 * 128, YPAD = (YBYTES + 127) / 128 * 128;
  static constexpr int STAGE = XPAD + YPAD;
  static constexpr int RED_TAP = 48 * CSV + 1;    // odd pitch per (kd, kh): the flush is conflict-free
  static constexpr int NRED = 9 * RED_TAP + 4 * CSV;   // aliases stage 0 after the last tile
  static constexpr size_t SMEM = NST * STAGE;
  static_assert(TD * TH == 3 * NWARP, "one 3-row group per warp");
  static_assert((PW % 8) == 1 && (XPLANE % 8) == 5, "kS2LaneTap assumes these bank-group pitches");
  static_assert(NRED * 4 <= STAGE, "reduction scratch fits a stage");
};

__device__ __forceinline__ void mbar_wait_parity(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "S2_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra S2_DONE;\n\t"
      "bra S2_WAIT;\n\t"
      "S2_DONE:\n\t"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}

// Warp roles: warps 0..3 run the FMA loop, one lane of warp 4 issues the tensor copies.  full[s]
// (1 arrival + the copies' bytes) hands a stage to the consumers, empty[s] (one arrival per
// consumer warp) hands it back; no CTA-wide barrier inside the tile loop.
// DW = true: the depthwise weight gradient dW[c][tap] += sum_o small[o, c] * big[2o - 1 + tap, c] on
// the same tiles (conv_tiled_dw.cu's stride-2 wgrad was the cp.async kernel's twin): blockIdx.y is
// the 4-channel chunk of both tensors, 3 float4 accumulators per lane.
template <int TWT, int CSV, bool BIAS, bool DW = false>
__global__ void __launch_bounds__(WgS2TmaShape<TWT, CSV>::THREADS)
    wgrad3_s2_tma_kernel(const S2Args A, int ntiles, const __grid_constant__ CUtensorMap xmap,
                         const __grid_constant__ CUtensorMap ymap, int x_merged, int y_merged) {
  using WS = WgS2TmaShape<TWT, CSV>;
  constexpr int TW = WS::TW, TH = WS::TH, PW = WS::PW, NST = WS::NST;
  extern __shared__ __align__(128) unsigned char base[];
  __shared__ __align__(8) unsigned long long full_bar[NST], empty_bar[NST];
  float* red = reinterpret_cast<float*>(base);

  static_assert(!DW || CSV == 1, "depthwise: one channel chunk");
  const int C4S = DW ? 1 : A.Cs / (4 * CSV);
  const int cic = DW ? (int)blockIdx.y : (int)blockIdx.y / C4S;   // cic: 4 big channels
  const int coc = DW ? (int)blockIdx.y : (int)blockIdx.y % C4S;   // coc: 4*CSV small channels
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned code = kS2LaneTap[lane];
  const int kd = code & 3, kh = (code >> 2) & 3, j = (code >> 4) & 3, role = code >> 6;
  const bool tap_lane = role == 0, active = role < 2;

  float2 acc[CSV][3][4][2];   // [cs chunk][kw][cb][cs pair]: FFMA2 accumulators
  float4 dacc[3];             // depthwise: [kw], one value per channel
#pragma unroll
  for (int t = 0; t < 3; ++t) dacc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 bsum[CSV];
#pragma unroll
  for (int v = 0; v < CSV; ++v) {
    bsum[v] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 3; ++t)
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c) acc[v][t][a][c] = make_float2(0.f, 0.f);
  }
  if (threadIdx.x == 0) {
    for (int b = 0; b < NST; ++b) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"((unsigned)__cvta_generic_to_shared(&full_bar[b])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(&empty_bar[b])),
                   "r"(WS::NWARP));
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (warp == WS::NWARP) {
    // ---- producer ---------------------------------------------------------------------------
    if (lane == 0) {
      int buf = 0, parity = 0;
#pragma unroll 1
      for (long long tile = blockIdx.x, k = 0; tile < ntiles; tile += gridDim.x, ++k) {
        if (k >= NST)    // the consumers have handed the stage back (its previous use: round k/NST - 1)
          mbar_wait_parity((unsigned)__cvta_generic_to_shared(&empty_bar[buf]), (unsigned)(parity ^ 1));
        unsigned b = (unsigned)tile;
        const int tw = (int)(b % (unsigned)A.tiles_w); b /= (unsigned)A.tiles_w;
        const int th = (int)(b % (unsigned)A.tiles_h); b /= (unsigned)A.tiles_h;
        const int td = (int)(b % (unsigned)A.tiles_d);
        const int n = (int)(b / (unsigned)A.tiles_d);
        const int w0 = tw * TW, h0 = th * TH, d0 = td * WS::TD;     // small coordinates
        const unsigned bar = (unsigned)__cvta_generic_to_shared(&full_bar[buf]);
        const unsigned dx = (unsigned)__cvta_generic_to_shared(base + buf * WS::STAGE);
        const unsigned dy = dx + WS::XPAD;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar),
                     "r"((unsigned)(WS::XBYTES + WS::YBYTES))
                     : "memory");
        if (x_merged)
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes "
              "[%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::"r"(dx),
              "l"(&xmap), "r"((2 * w0 - 1) * 4), "r"(2 * h0 - 1), "r"(2 * d0 - 1), "r"(n), "r"(bar)
              : "memory");
        else
          asm volatile(
              "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes "
              "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];\n" ::"r"(dx),
              "l"(&xmap), "r"(cic * 4), "r"(2 * w0 - 1), "r"(2 * h0 - 1), "r"(2 * d0 - 1), "r"(n), "r"(bar)
              : "memory");
        if (y_merged)
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes "
              "[%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::"r"(dy),
              "l"(&ymap), "r"(w0 * 4 * CSV), "r"(h0), "r"(d0), "r"(n), "r"(bar)
              : "memory");
        else
          asm volatile(
              "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes "
              "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];\n" ::"r"(dy),
              "l"(&ymap), "r"(coc * 4 * CSV), "r"(w0), "r"(h0), "r"(d0), "r"(n), "r"(bar)
              : "memory");
        if (++buf == NST) { buf = 0; parity ^= 1; }
      }
    }
  } else {
    // ---- consumers --------------------------------------------------------------------------
    const int row = warp * 3 + j;
    const int pd = row / TH, ph = row % TH;
    const int xoff = (2 * pd + kd) * WS::XPLANE + (2 * ph + kh) * PW, yoff = row * WS::YROW;
    int buf = 0, parity = 0;
#pragma unroll 1
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      mbar_wait_parity((unsigned)__cvta_generic_to_shared(&full_bar[buf]), (unsigned)parity);
      if (active) {
        const float4* xr = reinterpret_cast<const float4*>(base + buf * WS::STAGE) + xoff;
        const float4* yr = reinterpret_cast<const float4*>(base + buf * WS::STAGE + WS::XPAD) + yoff;
        float4 xa = xr[0];
#pragma unroll
        for (int w = 0; w < TW; ++w) {
          const float4 xb2 = xr[2 * w + 1];
          const float4 xc = xr[2 * w + 2];
          const float4 xv3[3] = {xa, xb2, xc};
#pragma unroll
          for (int v = 0; v < CSV; ++v) {
            const float4 g = yr[w * CSV + v];
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              if constexpr (DW) {
                dacc[kw].x = fmaf(xv3[kw].x, g.x, dacc[kw].x); dacc[kw].y = fmaf(xv3[kw].y, g.y, dacc[kw].y);
                dacc[kw].z = fmaf(xv3[kw].z, g.z, dacc[kw].z); dacc[kw].w = fmaf(xv3[kw].w, g.w, dacc[kw].w);
              } else {
                outer4(acc[v][kw], xv3[kw], g);
              }
            }
            if (BIAS) { bsum[v].x += g.x; bsum[v].y += g.y; bsum[v].z += g.z; bsum[v].w += g.w; }
          }
          xa = xc;
        }
      }
      __syncwarp();
      if (lane == 0)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(
                         (unsigned)__cvta_generic_to_shared(&empty_bar[buf]))
                     : "memory");
      if (++buf == NST) { buf = 0; parity ^= 1; }
    }
  }

  // ---- flush once per CTA (the ring is idle: every issued tile has been consumed) ----------------
  __syncthreads();
  for (int i = threadIdx.x; i < WS::NRED; i += WS::THREADS) red[i] = 0.f;
  __syncthreads();
  const int kdkh = kd * 3 + kh;
  if constexpr (DW) {
    // red[tap][4] -> dW[c][tap]
    if (warp < WS::NWARP && tap_lane) {
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        float* r = red + (kdkh * 3 + kw) * 4;
        atomicAdd(r + 0, dacc[kw].x); atomicAdd(r + 1, dacc[kw].y);
        atomicAdd(r + 2, dacc[kw].z); atomicAdd(r + 3, dacc[kw].w);
      }
    } else if (BIAS && warp < WS::NWARP && role == 1) {
      atomicAdd(&red[27 * 4 + 0], bsum[0].x); atomicAdd(&red[27 * 4 + 1], bsum[0].y);
      atomicAdd(&red[27 * 4 + 2], bsum[0].z); atomicAdd(&red[27 * 4 + 3], bsum[0].w);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 27 * 4; i += WS::THREADS)
      atomicAdd(A.dW + (long long)(cic * 4 + i % 4) * 27 + i / 4, red[i]);
    if (BIAS && A.dbias_small && threadIdx.x < 4)
      atomicAdd(A.dbias_small + cic * 4 + threadIdx.x, red[27 * 4 + threadIdx.x]);
    return;
  }
  if (warp < WS::NWARP) {
    if (tap_lane) {
#pragma unroll
      for (int v = 0; v < CSV; ++v)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c)
              atomicAdd(&red[kdkh * WS::RED_TAP + (kw * 4 + a) * (4 * CSV) + v * 4 + c],
                        (c & 1) ? acc[v][kw][a][c >> 1].y : acc[v][kw][a][c >> 1].x);
    } else if (BIAS && role == 1) {
#pragma unroll
      for (int v = 0; v < CSV; ++v) {
        atomicAdd(&red[9 * WS::RED_TAP + v * 4 + 0], bsum[v].x); atomicAdd(&red[9 * WS::RED_TAP + v * 4 + 1], bsum[v].y);
        atomicAdd(&red[9 * WS::RED_TAP + v * 4 + 2], bsum[v].z); atomicAdd(&red[9 * WS::RED_TAP + v * 4 + 3], bsum[v].w);
      }
    }
  }
  __syncthreads();
  // red[kd,kh][kw][cb][cs] -> dW[cs][cb][tap]
  for (int i = threadIdx.x; i < 27 * 16 * CSV; i += WS::THREADS) {
    const int cs = coc * 4 * CSV + i % (4 * CSV), cb = cic * 4 + (i / (4 * CSV)) % 4, t = i / (16 * CSV);
    atomicAdd(A.dW + ((long long)cs * A.Cb + cb) * 27 + t, red[(t / 3) * WS::RED_TAP + i - (t / 3) * 48 * CSV]);
  }
  if (BIAS && A.dbias_small && cic == 0 && threadIdx.x < 4 * CSV)
    atomicAdd(A.dbias_small + coc * 4 * CSV + threadIdx.x, red[9 * WS::RED_TAP + threadIdx.x]);
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

template <int TWT, int CSV, bool BIAS, bool DW = false>
static int launch_s2_wgrad_tma_b(S2Args A, cudaStream_t st) {
  using WS = WgS2TmaShape<TWT, CSV>;
  CUtensorMap xmap, ymap;
  memset(&xmap, 0, sizeof(xmap));
  memset(&ymap, 0, sizeof(ymap));
  int x_merged = 0, y_merged = 0;
  if (A.Cb == 4 && A.ld_big == 4 &&
      make_ndhwc_merged_map(&xmap, A.big, 4, A.Wb, A.Hb, A.Db, A.N, WS::PW, WS::PH, WS::PD))
    x_merged = 1;
  else if (!make_ndhwc_map(&xmap, A.big, A.Cb, A.Wb, A.Hb, A.Db, A.N, A.ld_big, WS::PW, WS::PH, WS::PD))
    return NAS3D_ERR_UNSUPPORTED;
  if (A.Cs == 4 * CSV && A.ld_small == A.Cs &&
      make_ndhwc_merged_map(&ymap, A.small, A.Cs, A.Ws, A.Hs, A.Ds, A.N, WS::YW, WS::TH, WS::TD))
    y_merged = 1;
  else if (!make_ndhwc_map(&ymap, A.small, A.Cs, A.Ws, A.Hs, A.Ds, A.N, A.ld_small, WS::YW, WS::TH, WS::TD,
                           4 * CSV))
    return NAS3D_ERR_UNSUPPORTED;
  A.tiles_w = (A.Ws + WS::TW - 1) / WS::TW;
  A.tiles_h = (A.Hs + WS::TH - 1) / WS::TH;
  A.tiles_d = (A.Ds + WS::TD - 1) / WS::TD;
  auto kern = wgrad3_s2_tma_kernel<TWT, CSV, BIAS, DW>;
  static int occ = 0;
  if (!occ) {
    NAS3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS::SMEM));
    NAS3D_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WS::THREADS, WS::SMEM));
    if (occ < 1) occ = 1;
  }
  const long long ntiles = (long long)A.N * A.tiles_w * A.tiles_h * A.tiles_d;
  const int pairs = DW ? A.Cb / 4 : (A.Cb / 4) * (A.Cs / (4 * CSV));
  long long gx = (long long)kNumSMs * occ / pairs;
  if (gx < 1) gx = 1;
  if (gx > ntiles) gx = ntiles;
  kern<<<dim3((unsigned)gx, pairs), WS::THREADS, WS::SMEM, st>>>(A, (int)ntiles, xmap, ymap, x_merged, y_merged);
  return launched(DW ? "dw_wgrad3_s2_tma" : (CSV == 2 ? "wgrad3_s2_tma_cs8" : "wgrad3_s2_tma"));
}

template <int TWT>
static int launch_s2_wgrad_tma(const S2Args& A, cudaStream_t st) {
  if (A.Cs % 8 == 0)
    return A.dbias_small ? launch_s2_wgrad_tma_b<TWT, 2, true>(A, st) : launch_s2_wgrad_tma_b<TWT, 2, false>(A, st);
  return A.dbias_small ? launch_s2_wgrad_tma_b<TWT, 1, true>(A, st) : launch_s2_wgrad_tma_b<TWT, 1, false>(A, st);
}

// depthwise stride-2 weight gradient on the TMA tile ring; NAS3D_ERR_UNSUPPORTED = not taken
int tiled_dw_s2_wgrad_tma(const S2Args& A, cudaStream_t st) {
  const bool ok = g_opt.s2_wgrad_tma && A.Cb == A.Cs && A.Cb % 4 == 0 && A.ld_big % 4 == 0 &&
                  A.ld_small % 4 == 0 && aligned16(A.big) && aligned16(A.small) && A.Db == 2 * A.Ds &&
                  A.Hb == 2 * A.Hs && A.Wb == 2 * A.Ws && A.Ws >= 2;
  if (!ok) return NAS3D_ERR_UNSUPPORTED;
  if (A.Ws <= 8)
    return A.dbias_small ? launch_s2_wgrad_tma_b<8, 1, true, true>(A, st) : launch_s2_wgrad_tma_b<8, 1, false, true>(A, st);
  return A.dbias_small ? launch_s2_wgrad_tma_b<16, 1, true, true>(A, st) : launch_s2_wgrad_tma_b<16, 1, false, true>(A, st);
}

int tiled_s2_wgrad(const S2Args& A, cudaStream_t st) {
  const bool ok = A.ld_big % 4 == 0 && A.ld_small % 4 == 0 && aligned16(A.big) && aligned16(A.small) &&
                  A.Db == 2 * A.Ds && A.Hb == 2 * A.Hs && A.Wb == 2 * A.Ws && A.Ws >= 2;
  if (!ok || A.Cb % 4 || A.Cs % 4 || A.Cb > 64 || A.Cs > 64) return NAS3D_ERR_UNSUPPORTED;
  if ((long long)A.Db * A.Hb * A.Wb * A.ld_big >= (1ll << 31)) return NAS3D_ERR_UNSUPPORTED;
  if (g_opt.s2_wgrad_tma) {
    const int rc = A.Ws <= 8 ? launch_s2_wgrad_tma<8>(A, st) : launch_s2_wgrad_tma<16>(A, st);
    if (rc != NAS3D_ERR_UNSUPPORTED) return rc;
  }
  if (A.Ws <= 8) return launch_s2_wgrad<8>(A, st);
  if (A.Ws <= 16) return launch_s2_wgrad<16>(A, st);
  return launch_s2_wgrad<32>(A, st);
}

}  // namespace nas3d
