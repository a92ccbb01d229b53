// Tiled direct 3x3x3 stride-1 convolution (dilation 1 or 2, Cin == Cout == C in {4,8,16}) on
// NDHWC fp32: the shapes that carry ~70 % of the network's FLOPs and bytes (C=4 @128^3,
// C=8 @64^3).
//
// Why CUDA-core FFMA and not tcgen05 here: with N = Cout = 4..8 a UMMA tile re-reads its A
// operand (128 voxels x K) from shared memory for 4..8 output columns only, so the tensor pipe
// is bounded by the smem operand bandwidth at ~1/8 of its peak, and 3xTF32 (needed for the
// 1e-3 gradient tolerance) divides that by 3 again - below what exact-fp32 FFMA delivers.
// (conv_umma.cu holds the tcgen05 path for the wide layers.)  See DESIGN.md "Kernels".
//
// One CTA = one output tile TD x TH x 32 voxels.  The input tile + halo is staged ONCE in
// shared memory (cp.async, zero-filled outside the volume = the conv's zero padding) in
// channel-chunk planes [C/4][PD][PH][PW] of float4 so that a warp's 32 lanes (consecutive w)
// read 512 contiguous bytes.  A thread owns MD x 4 (d x h) output voxels x 4 output channels:
// for every (kd,kw,cin-chunk) it loads the 3x4x4 weights once (broadcast LDS) and a column of
// 4+2*DIL input float4 per d, and issues 48 FFMA per loaded float4.
//
//   forward : out[o,co]  = bias[co] + sum_{tap,ci} x[o - pad + tap*dil, ci] * W[co][ci][tap]
//   dgrad   : dx[i,ci]   =           sum_{tap,co} dy[i - pad + (2-tap)*dil, co] * W[co][ci][tap]
// (same kernel: FLIP re-indexes the weights while they are staged).
//
//   wgrad   : dW[co][ci][tap] += sum_o dy[o,co] * x[o - pad + tap*dil, ci]
// lanes <-> taps: lane t < 27 owns the 4x4 (ci,co) block of tap t for the warp's voxels, so no
// cross-lane reduction is needed until the end; lane 27 accumulates the bias gradient.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "conv_tiled.h"

namespace nas3d {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

template <int C, int DIL, int HG, int DG, int MD>
struct TileShape {
  static constexpr int C4 = C / 4;
  static constexpr int TW = 32, TH = 4 * HG, TD = MD * DG;
  static constexpr int PW = TW + 2 * DIL, PH = TH + 2 * DIL, PD = TD + 2 * DIL;
  static constexpr int NR = 4 + 2 * DIL;
  static constexpr int THREADS = 32 * HG * DG * C4;
  static constexpr int PLANE = PD * PH * PW;                   // float4 per channel chunk
  static constexpr size_t SMEM = sizeof(float4) * PLANE * C4 + sizeof(float) * 27 * C * C;
};

// TMA = true: the haloed input tile is fetched by ONE elected thread with
// cp.async.bulk.tensor.5d (box {4 ch, PW, PH, PD, 1 sample} of the NDHWC tensor map, signed start
// coordinates, hardware zero-fill outside the volume = the conv padding) and lands on an mbarrier;
// no per-element address arithmetic is left in the instruction stream of this issue-bound kernel.
template <int C, int DIL, int HG, int DG, int MD, bool FLIP, bool TMA>
__global__ void __launch_bounds__(TileShape<C, DIL, HG, DG, MD>::THREADS)
    conv3_s1_kernel(const TiledArgs A, const __grid_constant__ CUtensorMap tmap) {
  using TS = TileShape<C, DIL, HG, DG, MD>;
  constexpr int C4 = TS::C4, PW = TS::PW, PH = TS::PH, PD = TS::PD, NR = TS::NR;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* tile = reinterpret_cast<float4*>(smem_raw);
  float* wsm = reinterpret_cast<float*>(smem_raw + sizeof(float4) * TS::PLANE * C4);
  __shared__ double sm_mom[2 * C];
  __shared__ __align__(8) unsigned long long tma_bar;
  if (threadIdx.x < 2 * C) sm_mom[threadIdx.x] = 0.0;

  int b = blockIdx.x;
  const int tw = b % A.tiles_w; b /= A.tiles_w;
  const int th = b % A.tiles_h; b /= A.tiles_h;
  const int td = b % A.tiles_d;
  const int n = b / A.tiles_d;
  const int w0 = tw * TS::TW, h0 = th * TS::TH, d0 = td * TS::TD;

  // ---- stage weights: wsm[tap][ci][co] ------------------------------------------------
  for (int i = threadIdx.x; i < 27 * C * C; i += TS::THREADS) {
    const int co = i % C, ci = (i / C) % C, tap = i / (C * C);
    // fwd : reduce over ci, produce co, tap as is
    // flip: reduce over (W's dim0), produce (W's dim1), tap mirrored
    const int gi = FLIP ? (ci * C + co) * 27 + (26 - tap) : (co * C + ci) * 27 + tap;
    wsm[i] = __ldg(A.w + gi);
  }
  // ---- stage the input tile + halo ------------------------------------------------------
  if constexpr (TMA) {
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&tma_bar);
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar),
                   "r"((unsigned)(sizeof(float4) * TS::PLANE * C4))
                   : "memory");
      if (C == 4 && A.tma_merged) {
        // dense C = 4: (W, C) merged into one 16*PW-byte inner box row - 18x fewer, 34x longer
        // TMA row requests than the 16-byte rows of the 5-D box (ncu r1e: 25 % of the warp time
        // of this kernel was spent waiting for the tile)
        const unsigned dst = (unsigned)__cvta_generic_to_shared(tile);
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes "
            "[%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::"r"(dst),
            "l"(&tmap), "r"((w0 - DIL) * 4), "r"(h0 - DIL), "r"(d0 - DIL), "r"(n), "r"(bar)
            : "memory");
      } else {
#pragma unroll
        for (int cc = 0; cc < C4; ++cc) {
          const unsigned dst = (unsigned)__cvta_generic_to_shared(tile + cc * TS::PLANE);
          asm volatile(
              "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes "
              "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];\n" ::"r"(dst),
              "l"(&tmap), "r"(cc * 4), "r"(w0 - DIL), "r"(h0 - DIL), "r"(d0 - DIL), "r"(n), "r"(bar)
              : "memory");
        }
      }
    }
    __syncthreads();     // barrier init visible to the waiters; weights staged
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "TMA_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
        "@p bra TMA_DONE;\n\t"
        "bra TMA_WAIT;\n\t"
        "TMA_DONE:\n\t"
        "}\n" ::"r"(bar)
        : "memory");
  } else {
    const float* xb = A.x + (long long)n * A.Dx * A.Hx * A.Wx * A.ldx;
    for (int i = threadIdx.x; i < TS::PLANE * C4; i += TS::THREADS) {
      const int cc = i % C4;
      int r = i / C4;
      const int pw = r % PW; r /= PW;
      const int ph = r % PH;
      const int pd = r / PH;
      const int gd = d0 - DIL + pd, gh = h0 - DIL + ph, gw = w0 - DIL + pw;
      const bool ok = gd >= 0 && gd < A.D && gh >= 0 && gh < A.H && gw >= 0 && gw < A.W;
      // 32-bit in-sample offset (host guarantees samples < 2^31 floats)
      const unsigned off = (unsigned)((((gd * A.Hx + gh) * A.Wx + gw) * A.xs) * A.ldx + cc * 4);
      cp_async16(&tile[cc * TS::PLANE + (pd * PH + ph) * PW + pw], ok ? xb + off : A.x, ok);
    }
    cp_async_wait_all();
    __syncthreads();
  }

  // ---- compute ----------------------------------------------------------------------------
  const int tx = threadIdx.x & 31;
  int ty = threadIdx.x >> 5;
  const int cg = ty % C4; ty /= C4;       // output-channel group (4 channels)
  const int hg = ty % HG;
  const int dg = ty / HG;

  float2 acc[MD][4][2];      // [d][h][channel pair]: FFMA2 accumulators (common.cuh)
#pragma unroll
  for (int a = 0; a < MD; ++a)
#pragma unroll
    for (int h = 0; h < 4; ++h)
#pragma unroll
      for (int c = 0; c < 2; ++c) acc[a][h][c] = make_float2(0.f, 0.f);

#pragma unroll 1
  for (int kd = 0; kd < 3; ++kd) {
#pragma unroll 1
    for (int kw = 0; kw < 3; ++kw) {
#pragma unroll 1
      for (int cc = 0; cc < C4; ++cc) {
        // weights of taps (kd, 0..2, kw), input channels cc*4..+3, my 4 output channels
        float4 wr[3][4];
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int ci = 0; ci < 4; ++ci)
            wr[kh][ci] = *reinterpret_cast<const float4*>(
                wsm + (((kd * 3 + kh) * 3 + kw) * C + cc * 4 + ci) * C + cg * 4);
#pragma unroll
        for (int od = 0; od < MD; ++od) {
          const float4* col = tile + cc * TS::PLANE +
                              ((dg * MD + od + kd * DIL) * PH + hg * 4) * PW + tx + kw * DIL;
          float4 xin[NR];
#pragma unroll
          for (int r = 0; r < NR; ++r) xin[r] = col[r * PW];
#pragma unroll
          for (int oh = 0; oh < 4; ++oh)
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) dot4x4(acc[od][oh], xin[oh + kh * DIL], wr[kh]);
        }
      }
    }
  }

  // ---- epilogue ---------------------------------------------------------------------------
  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (A.bias) bias4 = ldg4(A.bias + cg * 4);
  const int gw = w0 + tx;
  float ms[4] = {0.f, 0.f, 0.f, 0.f}, mq[4] = {0.f, 0.f, 0.f, 0.f};
  if (gw < A.W) {
    float* yb = A.y + (long long)n * A.Dy * A.Hy * A.Wy * A.ldy + cg * 4;
#pragma unroll
    for (int od = 0; od < MD; ++od) {
      const int gd = d0 + dg * MD + od;
      if (gd >= A.D) break;
#pragma unroll
      for (int oh = 0; oh < 4; ++oh) {
        const int gh = h0 + hg * 4 + oh;
        if (gh >= A.H) break;
        float* p = yb + (((long long)gd * A.ys * A.Hy + gh * A.ys) * A.Wy + gw * A.ys) * A.ldy;
        float4 v = make_float4(acc[od][oh][0].x + bias4.x, acc[od][oh][0].y + bias4.y,
                               acc[od][oh][1].x + bias4.z, acc[od][oh][1].y + bias4.w);
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
  if (A.moments) {      // uniform per launch
    warp_moments_add(sm_mom, cg * 4, ms, mq);
    __syncthreads();
    cta_moments_flush(sm_mom, A.moments, n, C, TS::THREADS);
  }
}

// -----------------------------------------------------------------------------------------
// wgrad:  dW[co][ci][tap] += sum_o dy[o,co] * x[o - DIL + tap*DIL, ci]   (lattice coordinates)
//
// The CTA stages the x tile (+halo) and the dy tile in shared memory.  Inside a warp, lane
// l < 27 is (kdkh = l / 3, j = l % 3): it owns the three taps (kd, kh, kw = 0..2) for tile row
// (3*i + j) and WALKS ALONG W with a sliding window of 2*DIL+1 x-values, so each step costs one
// new 128-bit x load + one dy load for 3 taps x 16 = 48 FFMA.  All partial sums stay in the
// lane's registers until the end of the tile (no cross-lane reduction in the hot loop).
// Lanes 27..29 ride along and accumulate the bias gradient (sum of dy) of rows j = 0..2.
// For C > 4 the (ci-chunk, co-chunk) pairs are looped over the same staged tile.
// -----------------------------------------------------------------------------------------
template <int DIL, int TH, int TD, int NWARP, int TWT = 32>
struct WgShape {
  static constexpr int TW = TWT;      // 32 | 16 | 8: narrow volumes (low-resolution wide layers)
  static constexpr int PW = TW + 2 * DIL, PH = TH + 2 * DIL, PD = TD + 2 * DIL;
  // Row / plane / dy-row pitches (in float4) found by exhaustive search so that every 8-lane
  // phase of the two LDS.128 per step hits 8 distinct 16-byte bank groups (8 wavefronts per
  // step = the minimum; the naive layout needs 26).
  static constexpr int PWP = PW | 1;
  static constexpr int PLANE_PAD = DIL == 1 ? (TH == 12 ? 5 : 7) : (TH == 12 ? 0 : 2);
  static constexpr int XPLANE = PH * PWP + PLANE_PAD;       // pitch of one d-plane
  static constexpr int XTILE = PD * XPLANE + 8;
  static constexpr int YP = TW + 1;                          // dy row pitch
  static constexpr int YTILE = TD * TH * YP + 8;
  static constexpr int THREADS = 32 * NWARP;
  static constexpr int WN = 2 * DIL + 1;              // sliding window length
  static constexpr int ROWS = TD * TH;
  static constexpr int NGROUPS = ROWS / 3;            // 3 adjacent rows per warp step
  static_assert(ROWS % 3 == 0 && NGROUPS % NWARP == 0, "tile rows must split into 3-row groups per warp");
  static_assert(TH == 12 || TH == 6, "bank-conflict-free pitches are tabulated for TH = 12 / 6");
  static constexpr int NBUF = 2;                      // double-buffered tiles (prefetch of tile i+1)
  static constexpr size_t SMEM = sizeof(float4) * NBUF * (XTILE + YTILE) + sizeof(float) * (27 * 16 + 4);
};

// PERSISTENT: blockIdx.x strides over the tiles, blockIdx.y = (ci-chunk, co-chunk) pair; the
// 48 partial sums of a lane live in registers across all tiles of the CTA and are flushed once.
// Staging cost matters here (it was 27 % of the instruction stream, ncu r1e): the x tile is
// staged ROW-wise (a warp owns a (d,h) row: one 64-bit row pointer, then 32-bit per-lane offsets)
// and the dy tile with 32-bit in-sample offsets (host guarantees samples < 2^31 floats).
// BIAS = false drops the bias-gradient adds from the hot loop (convs followed by GroupNorm get
// their bias gradient analytically, engine.py).
// The tiles are DOUBLE-BUFFERED: the cp.async copies of the CTA's next tile are issued before the
// FMA loop of the current one (one __syncthreads per tile), so the FMA pipe no longer idles while
// a whole CTA waits for its tile (ncu r1e: pipe 58 % busy with staging and compute serialised).
template <int DIL, int TH, int TD, int NWARP, int TWT, bool BIAS>
__global__ void __launch_bounds__(WgShape<DIL, TH, TD, NWARP, TWT>::THREADS)
    wgrad3_s1_kernel(const WgradArgs A, int C, int ntiles) {
  using WS = WgShape<DIL, TH, TD, NWARP, TWT>;
  constexpr int PW = WS::PW, PWP = WS::PWP, PH = WS::PH, TW = WS::TW, WN = WS::WN;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* const tiles = reinterpret_cast<float4*>(smem_raw);   // [NBUF][XTILE + YTILE]
  float* red = reinterpret_cast<float*>(tiles + WS::NBUF * (WS::XTILE + WS::YTILE));   // [27][4][4] + [4]

  const int C4 = C / 4;
  const int cic = blockIdx.y / C4, coc = blockIdx.y % C4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool tap_lane = lane < 27;
  const int kdkh = tap_lane ? lane / 3 : 0;
  const int j = lane < 30 ? lane % 3 : 0;
  const bool active = lane < 30;
  const int kd = kdkh / 3, kh = kdkh % 3;

  float2 acc[3][4][2];   // [kw][ci][co pair]: FFMA2 accumulators
#pragma unroll
  for (int t = 0; t < 3; ++t)
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 2; ++c) acc[t][a][c] = make_float2(0.f, 0.f);
  float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int i = threadIdx.x; i < 27 * 16 + 4; i += WS::THREADS) red[i] = 0.f;

  // ---- staging: THREAD-COLUMN scheme.  A thread owns one w column of the haloed x tile (and one of
  // the dy tile) for the whole kernel and walks down the (d,h) rows RPI rows at a time, so the
  // per-copy instruction stream is ~18 integer ops (ncu r1e: with a per-element index
  // decomposition the staging code issued as many instructions as the FMA loop and the two do
  // not overlap in the issue slot).  32-bit in-sample offsets (host guarantees < 2^31 floats).
  constexpr int XROWS = WS::PD * PH, RPI_X = WS::THREADS / PW, XITER = (XROWS + RPI_X - 1) / RPI_X;
  constexpr int RPI_Y = WS::THREADS / TW, YITER = (WS::ROWS + RPI_Y - 1) / RPI_Y;
  const int my_pw = threadIdx.x % PW, my_xr0 = threadIdx.x / PW;
  const bool x_thread = my_xr0 < RPI_X;
  const int my_yw = threadIdx.x % TW, my_yr0 = threadIdx.x / TW;
  const int xstep = A.xs * A.ldx;                         // floats between lattice neighbours along w
  const unsigned xrow = (unsigned)(A.xs * A.Wx * A.ldx);   // ... along h (lattice rows)
  const unsigned yrow = (unsigned)(A.ys * A.Wy * A.ldy);
  // issue the cp.async copies of one tile into buffer `buf` (one commit group per thread)
  auto stage = [&](int tile, int buf) {
    float4* xt = tiles + buf * (WS::XTILE + WS::YTILE);
    float4* yt = xt + WS::XTILE;
    unsigned b = (unsigned)tile;
    const unsigned tw = b % (unsigned)A.tiles_w; b /= (unsigned)A.tiles_w;
    const unsigned th = b % (unsigned)A.tiles_h; b /= (unsigned)A.tiles_h;
    const unsigned td = b % (unsigned)A.tiles_d;
    const unsigned n = b / (unsigned)A.tiles_d;
    const int w0 = (int)tw * TW, h0 = (int)th * TH, d0 = (int)td * TD;
    {
      const int gw = w0 - DIL + my_pw;
      const bool cok = (unsigned)gw < (unsigned)A.W;
      const float* xcol = A.x + (long long)n * A.Dx * A.Hx * A.Wx * A.ldx + cic * 4 + gw * xstep;
#pragma unroll
      for (int k = 0; k < XITER; ++k) {
        const int r = my_xr0 + k * RPI_X;
        if (x_thread && r < XROWS) {
          const int pd = r / PH, ph = r - pd * PH;
          const int gd = d0 - DIL + pd, gh = h0 - DIL + ph;
          const bool ok = cok && (unsigned)gd < (unsigned)A.D && (unsigned)gh < (unsigned)A.H;
          cp_async16(xt + pd * WS::XPLANE + ph * PWP + my_pw,
                     ok ? xcol + (unsigned)(gd * A.Hx + gh) * xrow : A.x, ok);
        }
      }
    }
    {
      const int gw = w0 + my_yw;
      const bool cok = gw < A.W;                           // outside: dy = 0 contributes nothing
      const float* ycol = A.dy + (long long)n * A.Dy * A.Hy * A.Wy * A.ldy + coc * 4 + gw * (A.ys * A.ldy);
#pragma unroll
      for (int k = 0; k < YITER; ++k) {
        const int r = my_yr0 + k * RPI_Y;
        if (r < WS::ROWS) {
          const int pd = r / TH, ph = r - pd * TH;
          const int gd = d0 + pd, gh = h0 + ph;
          const bool ok = cok && gd < A.D && gh < A.H;
          cp_async16(yt + r * WS::YP + my_yw, ok ? ycol + (unsigned)(gd * A.Hy + gh) * yrow : A.dy, ok);
        }
      }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };

  if ((int)blockIdx.x < ntiles) stage(blockIdx.x, 0);
  int buf = 0;
#pragma unroll 1
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");   // my copies of `tile` have landed
    __syncthreads();   // ... everyone's have, and everyone is done with the other buffer
    if (tile + (int)gridDim.x < ntiles) stage(tile + gridDim.x, buf ^ 1);   // in flight during the FMAs
    const float4* xt = tiles + buf * (WS::XTILE + WS::YTILE);
    const float4* yt = xt + WS::XTILE;

#pragma unroll 1
    for (int grp = warp; grp < WS::NGROUPS; grp += NWARP) {
      if (!active) continue;
      const int row = grp * 3 + j;
      const int pd = row / TH, ph = row % TH;
      const float4* xr = xt + (pd + kd * DIL) * WS::XPLANE + (ph + kh * DIL) * PWP;
      const float4* yr = yt + (pd * TH + ph) * WS::YP;
      // circular window: slot q holds xr[wb + q] at the top of every WN-step block
      float4 win[WN];
#pragma unroll
      for (int q = 0; q < WN - 1; ++q) win[q] = xr[q];
      constexpr int NFULL = TW / WN;
#pragma unroll 2
      for (int blk = 0; blk < NFULL; ++blk) {
        const float4* xq = xr + blk * WN;
        const float4* yq = yr + blk * WN;
#pragma unroll
        for (int u = 0; u < WN; ++u) {
          win[(u + WN - 1) % WN] = xq[u + WN - 1];
          const float4 g = yq[u];
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) outer4(acc[kw], win[(u + kw * DIL) % WN], g);
          if (BIAS) { bsum.x += g.x; bsum.y += g.y; bsum.z += g.z; bsum.w += g.w; }
        }
      }
      // tail: TW % WN steps
      {
        const float4* xq = xr + NFULL * WN;
        const float4* yq = yr + NFULL * WN;
#pragma unroll
        for (int u = 0; u < TW % WN; ++u) {
          win[(u + WN - 1) % WN] = xq[u + WN - 1];
          const float4 g = yq[u];
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) outer4(acc[kw], win[(u + kw * DIL) % WN], g);
          if (BIAS) { bsum.x += g.x; bsum.y += g.y; bsum.z += g.z; bsum.w += g.w; }
        }
      }
    }
  }

  // flush once per CTA
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
  // red[tap][ci][co] -> dW[co][ci][tap]
  for (int i = threadIdx.x; i < 27 * 16; i += WS::THREADS) {
    const int co = coc * 4 + i % 4, ci = cic * 4 + (i / 4) % 4, t = i / 16;
    atomicAdd(A.dW + (co * C + ci) * 27 + t, red[i]);
  }
  if (BIAS && A.dbias && cic == 0 && threadIdx.x < 4)
    atomicAdd(A.dbias + coc * 4 + threadIdx.x, red[27 * 16 + threadIdx.x]);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return g_opt.tma ? fn : nullptr;
}

// 5-D map of an NDHWC fp32 tensor {C, W, H, D, N} with voxel pitch ld, box {bc, bw, bh, bd, 1}
bool make_ndhwc_map(CUtensorMap* m, const float* base, int C, int W, int H, int D, int N,
                    int ld, int bw, int bh, int bd, int bc) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  const cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  const cuuint64_t strides[4] = {(cuuint64_t)ld * 4, (cuuint64_t)W * ld * 4, (cuuint64_t)H * W * ld * 4,
                                 (cuuint64_t)D * H * W * ld * 4};
  const cuuint32_t box[5] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bd, 1};
  const cuuint32_t es[5] = {1, 1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(base), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// 4-D map {W*c, H, D, N} of a DENSE c-channel NDHWC tensor (ld == c), box {c*bw, bh, bd, 1}
bool make_ndhwc_merged_map(CUtensorMap* m, const float* base, int c, int W, int H, int D, int N,
                           int bw, int bh, int bd) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc || !g_opt.tma_merged || c * bw > 256) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)W * c, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  const cuuint64_t strides[3] = {(cuuint64_t)W * c * 4, (cuuint64_t)H * W * c * 4,
                                 (cuuint64_t)D * H * W * c * 4};
  const cuuint32_t box[4] = {(cuuint32_t)(c * bw), (cuuint32_t)bh, (cuuint32_t)bd, 1};
  const cuuint32_t es[4] = {1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int C, int DIL, int HG, int DG, int MD, bool FLIP>
static int launch_conv3(const TiledArgs& A0, cudaStream_t st) {
  using TS = TileShape<C, DIL, HG, DG, MD>;
  TiledArgs A = A0;
  A.tiles_w = (A.W + TS::TW - 1) / TS::TW;
  A.tiles_h = (A.H + TS::TH - 1) / TS::TH;
  A.tiles_d = (A.D + TS::TD - 1) / TS::TD;
  auto kern_tma = conv3_s1_kernel<C, DIL, HG, DG, MD, FLIP, true>;
  auto kern_cp = conv3_s1_kernel<C, DIL, HG, DG, MD, FLIP, false>;
  static bool attr_done = false;
  if (!attr_done) {
    NAS3D_CUDA(cudaFuncSetAttribute(kern_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TS::SMEM));
    NAS3D_CUDA(cudaFuncSetAttribute(kern_cp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TS::SMEM));
    attr_done = true;
  }
  const long long blocks = (long long)A.N * A.tiles_w * A.tiles_h * A.tiles_d;
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  A.tma_merged = 0;
  bool tma = false;
  if (C == 4 && A.xs == 1 && A.ldx == 4 &&
      make_ndhwc_merged_map(&tmap, A.x, 4, A.W, A.H, A.D, A.N, TS::PW, TS::PH, TS::PD)) {
    tma = true;
    A.tma_merged = 1;
  } else {
    tma = A.xs == 1 && (TS::PLANE % 8 == 0 || TS::C4 == 1) &&
          make_ndhwc_map(&tmap, A.x, C, A.W, A.H, A.D, A.N, A.ldx, TS::PW, TS::PH, TS::PD);
  }
  if (tma) kern_tma<<<(unsigned)blocks, TS::THREADS, TS::SMEM, st>>>(A, tmap);
  else kern_cp<<<(unsigned)blocks, TS::THREADS, TS::SMEM, st>>>(A, tmap);
  return launched(!tma ? "conv3_s1_cpasync" : (A.tma_merged ? "conv3_s1_tma_merged" : "conv3_s1_tma5d"));
}

template <int DIL, int TH, int TD, int NWARP, int TWT, bool BIAS>
static int launch_wgrad3_b(const WgradArgs& A0, int C, cudaStream_t st) {
  using WS = WgShape<DIL, TH, TD, NWARP, TWT>;
  WgradArgs A = A0;
  A.tiles_w = (A.W + WS::TW - 1) / WS::TW;
  A.tiles_h = (A.H + TH - 1) / TH;
  A.tiles_d = (A.D + TD - 1) / TD;
  auto kern = wgrad3_s1_kernel<DIL, TH, TD, NWARP, TWT, BIAS>;
  static int occ = 0;
  if (!occ) {
    NAS3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS::SMEM));
    NAS3D_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WS::THREADS, WS::SMEM));
    if (occ < 1) occ = 1;
  }
  const long long ntiles = (long long)A.N * A.tiles_w * A.tiles_h * A.tiles_d;
  const int pairs = (C / 4) * (C / 4);
  long long gx = (long long)kNumSMs * occ / pairs;
  if (gx < 1) gx = 1;
  if (gx > ntiles) gx = ntiles;
  kern<<<dim3((unsigned)gx, pairs), WS::THREADS, WS::SMEM, st>>>(A, C, (int)ntiles);
  return launched("wgrad3_s1");
}

template <int DIL, int TH, int TD, int NWARP, int TWT>
static int launch_wgrad3(const WgradArgs& A, int C, cudaStream_t st) {
  return A.dbias ? launch_wgrad3_b<DIL, TH, TD, NWARP, TWT, true>(A, C, st)
                 : launch_wgrad3_b<DIL, TH, TD, NWARP, TWT, false>(A, C, st);
}

__global__ void __launch_bounds__(256)
    fill_channels_kernel(float* __restrict__ y, const float* __restrict__ bias, long long total,
                         int C4, int ld) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long v = i / C4;
    const int c = (int)(i - v * C4) * 4;
    st4(y + v * ld + c, ldg4(bias + c));
  }
}

int fill_channels(float* y, const float* bias, long long nvox, int C, int ld, cudaStream_t st) {
  NAS3D_REQUIRE(C % 4 == 0 && ld % 4 == 0 && aligned16(y) && aligned16(bias), "fill_channels: alignment");
  const long long total = nvox * (C / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  fill_channels_kernel<<<(unsigned)blocks, 256, 0, st>>>(y, bias, total, C / 4, ld);
  return launched("fill_channels");
}

// returns NAS3D_ERR_UNSUPPORTED (without setting an error) when the shape is not covered
int tiled_conv3_s1(int C, int dil, bool flip, const TiledArgs& A, cudaStream_t st) {
  if (A.W < 8 || A.ldx % 4 || A.ldy % 4 || !aligned16(A.x) || !aligned16(A.y)) return NAS3D_ERR_UNSUPPORTED;
  if ((long long)A.Dx * A.Hx * A.Wx * A.ldx >= (1ll << 31)) return NAS3D_ERR_UNSUPPORTED;   // 32-bit staging offsets
#define NAS3D_CASE(CC, DD, HG, DG, MD)                                          \
  if (C == CC && dil == DD)                                                      \
    return flip ? launch_conv3<CC, DD, HG, DG, MD, true>(A, st)                  \
                : launch_conv3<CC, DD, HG, DG, MD, false>(A, st);
  // (C, dil, HG, DG, MD): tile = 32 x 4*HG x MD*DG voxels.  Smaller tiles than round 1's where they
  // let another CTA share the SM (measured, tools/conv_micro.py, 8 x 128^3 / 8 x 64^3, fwd / dgrad us):
  // C4 dil 2 (4,2)->(2,2): 346 -> 315 / 312 -> 294; C8 dil 2 (2,2)->(1,2), one CTA per SM before:
  // 245 -> 202 / 228 -> 185; C8 dil 1 (2,2)->(1,2): 174 -> 164 / 167 -> 156; C16 dil 1 (2,1)->(1,1) at
  // 8 x 32^3: 100 -> 90 / 93 -> 83 (dil 2: slower, stays); C4 dil 1 stays (4,2).
  NAS3D_CASE(4, 1, 4, 2, 2)
  NAS3D_CASE(4, 2, 2, 2, 2)
  NAS3D_CASE(8, 1, 1, 2, 2)
  NAS3D_CASE(8, 2, 1, 2, 2)
  NAS3D_CASE(16, 1, 1, 1, 2)
  NAS3D_CASE(16, 2, 2, 1, 2)
#undef NAS3D_CASE
  return NAS3D_ERR_UNSUPPORTED;
}

// =========================================================================================
// stride-1 weight gradient on a TMA tile ring (same structure as wgrad3_s2_tma_kernel,
// conv_tiled_s2.cu): 16 FMA warps + 1 producer warp, the haloed x tile and the dy tile of a
// 32 x 12 x 4 voxel block arrive as two tensor copies (OOB zero-fill = the conv padding) on
// full[s], the FMA warps hand the stage back on empty[s]; no staging instructions and no CTA
// barrier in the tile loop.  The TMA tile is dense, so its pitches are box extents: the box is one
// row (dil 1: 15 x 34) / one column (dil 2: 16 x 37) larger than needed, and the lanes' (kd, kh,
// row) assignment is permuted so that every quarter-warp of the x LDS.128 hits 8 distinct bank
// groups or the same address (exhaustive search, as for the stride-2 kernel).
// byte = kd | kh << 2 | j << 4 | role << 6   (role 0: tap lane, 1: bias lane, 2: idle)
// =========================================================================================
#define S1T(kd, kh, j) ((kd) | (kh) << 2 | (j) << 4)
__constant__ unsigned char kS1LaneTapD1[32] = {
    S1T(0, 1, 0), S1T(0, 1, 2), S1T(0, 2, 1), S1T(0, 1, 1), S1T(0, 0, 2), S1T(0, 2, 0), S1T(0, 2, 2), S1T(0, 0, 1),
    S1T(1, 1, 2), S1T(0, 0, 0), S1T(2, 1, 2), S1T(2, 2, 1), S1T(1, 0, 0), S1T(1, 2, 1), S1T(0, 0, 0) | 64, S1T(2, 2, 1) | 64,
    S1T(1, 0, 1), S1T(2, 1, 0), S1T(2, 0, 1), S1T(1, 0, 2), S1T(2, 2, 2), S1T(1, 2, 0), S1T(1, 1, 1), S1T(1, 1, 0),
    S1T(1, 2, 2), S1T(2, 0, 0), S1T(2, 1, 1), S1T(2, 0, 2), S1T(2, 2, 0), S1T(1, 2, 2) | 64, 128, 128};
__constant__ unsigned char kS1LaneTapD2[32] = {
    S1T(0, 1, 0), S1T(0, 1, 2), S1T(2, 0, 1), S1T(0, 1, 1), S1T(0, 0, 2), S1T(0, 2, 0), S1T(0, 1, 0) | 64, S1T(2, 0, 1) | 64,
    S1T(1, 1, 2), S1T(0, 0, 0), S1T(2, 2, 1), S1T(2, 1, 1), S1T(2, 2, 2), S1T(1, 2, 0), S1T(0, 0, 1), S1T(1, 1, 2) | 64,
    S1T(1, 0, 1), S1T(0, 2, 1), S1T(2, 0, 0), S1T(1, 0, 2), S1T(0, 2, 2), S1T(1, 1, 1), S1T(1, 1, 0), 128,
    S1T(1, 2, 2), S1T(2, 1, 0), S1T(2, 1, 2), S1T(2, 0, 2), S1T(1, 0, 0), S1T(1, 2, 1), S1T(2, 2, 0), 128};
#undef S1T

template <int DIL, int CSV>
struct WgS1TmaShape {
  // 8 output channels per lane need ~140 registers: half the warps (and a tile half as deep) then
  static constexpr int TW = 32, TH = 12, TD = CSV == 2 ? 2 : 4, NWARP = TD * TH / 3;   // consumer warps; + 1 producer warp
  static constexpr int THREADS = 32 * (NWARP + 1);
  static constexpr int PD = TD + 2 * DIL;
  static constexpr int PH = TH + 2 * DIL + (DIL == 1 ? 1 : 0);    // box rows (one spare for dil 1)
  static constexpr int PW = TW + 2 * DIL + (DIL == 2 ? 1 : 0);    // box columns (one spare for dil 2)
  static constexpr int XPLANE = PH * PW;
  static constexpr int XBYTES = PD * XPLANE * 16;
  // dy box one column wider than the tile: with a row pitch of 32 float4 the 3 dy rows a warp reads
  // share a bank group and every dy LDS.128 took 12 wavefronts instead of 4 (ncu r8c)
  static constexpr int YW = TW + 1;
  static constexpr int YROW = YW * CSV;                            // float4 per dy row
  static constexpr int YBYTES = TD * TH * YROW * 16;
  static constexpr int XPAD = (XBYTES + 127) / 128 * 128, YPAD = (YBYTES + 127) / 128 * 128;
  static constexpr int STAGE = XPAD + YPAD;
  static constexpr int NST = 3 * STAGE <= 226 * 1024 ? 3 : 2;
  static constexpr int RED_TAP = 48 * CSV + 1;
  static constexpr int NRED = 9 * RED_TAP + 4 * CSV;               // aliases stage 0 after the last tile
  static constexpr size_t SMEM = NST * STAGE;
  static constexpr int WN = 2 * DIL + 1;
  static_assert(TD * TH == 3 * NWARP, "one 3-row group per warp");
  static_assert(NST * STAGE <= 227 * 1024, "tile ring fits shared memory");
  static_assert(DIL == 1 ? (PW % 8 == 2 && XPLANE % 8 == 6) : (PW % 8 == 5 && XPLANE % 8 == 0),
                "kS1LaneTap tables assume these bank-group pitches");
};

__device__ __forceinline__ void s1_mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "S1_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra S1_DONE;\n\t"
      "bra S1_WAIT;\n\t"
      "S1_DONE:\n\t"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}

template <int DIL, int CSV, bool BIAS>
__global__ void __launch_bounds__(WgS1TmaShape<DIL, CSV>::THREADS)
    wgrad3_s1_tma_kernel(const WgradArgs A, int C, int ntiles, const __grid_constant__ CUtensorMap xmap,
                         const __grid_constant__ CUtensorMap ymap, int x_merged, int y_merged) {
  using WS = WgS1TmaShape<DIL, CSV>;
  constexpr int TW = WS::TW, TH = WS::TH, PW = WS::PW, NST = WS::NST, WN = WS::WN;
  extern __shared__ __align__(128) unsigned char base[];
  __shared__ __align__(8) unsigned long long full_bar[NST], empty_bar[NST];
  float* red = reinterpret_cast<float*>(base);

  const int C4S = C / (4 * CSV);
  const int cic = blockIdx.y / C4S, coc = blockIdx.y % C4S;   // cic: 4 input channels, coc: 4*CSV output
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned code = DIL == 1 ? kS1LaneTapD1[lane] : kS1LaneTapD2[lane];
  const int kd = code & 3, kh = (code >> 2) & 3, j = (code >> 4) & 3, role = code >> 6;
  const bool tap_lane = role == 0, active = role < 2;

  float2 acc[CSV][3][4][2];   // [co chunk][kw][ci][co pair]: FFMA2 accumulators
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
        if (k >= NST)
          s1_mbar_wait((unsigned)__cvta_generic_to_shared(&empty_bar[buf]), (unsigned)(parity ^ 1));
        unsigned b = (unsigned)tile;
        const int tw = (int)(b % (unsigned)A.tiles_w); b /= (unsigned)A.tiles_w;
        const int th = (int)(b % (unsigned)A.tiles_h); b /= (unsigned)A.tiles_h;
        const int td = (int)(b % (unsigned)A.tiles_d);
        const int n = (int)(b / (unsigned)A.tiles_d);
        const int w0 = tw * TW, h0 = th * TH, d0 = td * WS::TD;
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
              "l"(&xmap), "r"((w0 - DIL) * 4), "r"(h0 - DIL), "r"(d0 - DIL), "r"(n), "r"(bar)
              : "memory");
        else
          asm volatile(
              "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes "
              "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];\n" ::"r"(dx),
              "l"(&xmap), "r"(cic * 4), "r"(w0 - DIL), "r"(h0 - DIL), "r"(d0 - DIL), "r"(n), "r"(bar)
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
    const int xoff = (pd + kd * DIL) * WS::XPLANE + (ph + kh * DIL) * PW, yoff = row * WS::YROW;
    int buf = 0, parity = 0;
#pragma unroll 1
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      s1_mbar_wait((unsigned)__cvta_generic_to_shared(&full_bar[buf]), (unsigned)parity);
      if (active) {
        const float4* xr = reinterpret_cast<const float4*>(base + buf * WS::STAGE) + xoff;
        const float4* yr = reinterpret_cast<const float4*>(base + buf * WS::STAGE + WS::XPAD) + yoff;
        // circular window: slot q holds xr[wb + q] at the top of every WN-step block
        float4 win[WN];
#pragma unroll
        for (int q = 0; q < WN - 1; ++q) win[q] = xr[q];
        constexpr int NFULL = TW / WN;
#pragma unroll 2
        for (int blk = 0; blk < NFULL; ++blk) {
          const float4* xq = xr + blk * WN;
          const float4* yq = yr + blk * WN * CSV;
#pragma unroll
          for (int u = 0; u < WN; ++u) {
            win[(u + WN - 1) % WN] = xq[u + WN - 1];
#pragma unroll
            for (int v = 0; v < CSV; ++v) {
              const float4 g = yq[u * CSV + v];
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) outer4(acc[v][kw], win[(u + kw * DIL) % WN], g);
              if (BIAS) { bsum[v].x += g.x; bsum[v].y += g.y; bsum[v].z += g.z; bsum[v].w += g.w; }
            }
          }
        }
        {
          const float4* xq = xr + NFULL * WN;
          const float4* yq = yr + NFULL * WN * CSV;
#pragma unroll
          for (int u = 0; u < TW % WN; ++u) {
            win[(u + WN - 1) % WN] = xq[u + WN - 1];
#pragma unroll
            for (int v = 0; v < CSV; ++v) {
              const float4 g = yq[u * CSV + v];
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) outer4(acc[v][kw], win[(u + kw * DIL) % WN], g);
              if (BIAS) { bsum[v].x += g.x; bsum[v].y += g.y; bsum[v].z += g.z; bsum[v].w += g.w; }
            }
          }
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
  // red[kd,kh][kw][ci][co] -> dW[co][ci][tap]
  for (int i = threadIdx.x; i < 27 * 16 * CSV; i += WS::THREADS) {
    const int co = coc * 4 * CSV + i % (4 * CSV), ci = cic * 4 + (i / (4 * CSV)) % 4, t = i / (16 * CSV);
    atomicAdd(A.dW + ((long long)co * C + ci) * 27 + t, red[(t / 3) * WS::RED_TAP + i - (t / 3) * 48 * CSV]);
  }
  if (BIAS && A.dbias && cic == 0 && threadIdx.x < 4 * CSV)
    atomicAdd(A.dbias + coc * 4 * CSV + threadIdx.x, red[9 * WS::RED_TAP + threadIdx.x]);
}

template <int DIL, int CSV, bool BIAS>
static int launch_wgrad3_tma_b(WgradArgs A, int C, cudaStream_t st) {
  using WS = WgS1TmaShape<DIL, CSV>;
  CUtensorMap xmap, ymap;
  memset(&xmap, 0, sizeof(xmap));
  memset(&ymap, 0, sizeof(ymap));
  int x_merged = 0, y_merged = 0;
  if (C == 4 && A.ldx == 4 && make_ndhwc_merged_map(&xmap, A.x, 4, A.W, A.H, A.D, A.N, WS::PW, WS::PH, WS::PD))
    x_merged = 1;
  else if (!make_ndhwc_map(&xmap, A.x, C, A.W, A.H, A.D, A.N, A.ldx, WS::PW, WS::PH, WS::PD))
    return NAS3D_ERR_UNSUPPORTED;
  if (C == 4 * CSV && A.ldy == C &&
      make_ndhwc_merged_map(&ymap, A.dy, C, A.W, A.H, A.D, A.N, WS::YW, WS::TH, WS::TD))
    y_merged = 1;
  else if (!make_ndhwc_map(&ymap, A.dy, C, A.W, A.H, A.D, A.N, A.ldy, WS::YW, WS::TH, WS::TD, 4 * CSV))
    return NAS3D_ERR_UNSUPPORTED;
  A.tiles_w = (A.W + WS::TW - 1) / WS::TW;
  A.tiles_h = (A.H + WS::TH - 1) / WS::TH;
  A.tiles_d = (A.D + WS::TD - 1) / WS::TD;
  auto kern = wgrad3_s1_tma_kernel<DIL, CSV, BIAS>;
  static bool attr_done = false;
  if (!attr_done) {
    NAS3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS::SMEM));
    attr_done = true;
  }
  const long long ntiles = (long long)A.N * A.tiles_w * A.tiles_h * A.tiles_d;
  const int pairs = (C / 4) * (C / (4 * CSV));
  long long gx = (long long)kNumSMs / pairs;           // one CTA per SM (the ring takes the shared memory)
  if (gx < 1) gx = 1;
  if (gx > ntiles) gx = ntiles;
  kern<<<dim3((unsigned)gx, pairs), WS::THREADS, WS::SMEM, st>>>(A, C, (int)ntiles, xmap, ymap, x_merged, y_merged);
  return launched(CSV == 2 ? "wgrad3_s1_tma_cs8" : "wgrad3_s1_tma");
}

// NAS3D_ERR_UNSUPPORTED = not taken (the cp.async kernel serves the shape)
static int launch_wgrad3_tma(int C, int dil, const WgradArgs& A, cudaStream_t st) {
  if (!g_opt.s1_wgrad_tma || A.xs != 1 || A.ys != 1 || A.W < 32) return NAS3D_ERR_UNSUPPORTED;
  if (C == 4) {
    if (dil == 1) return A.dbias ? launch_wgrad3_tma_b<1, 1, true>(A, C, st) : launch_wgrad3_tma_b<1, 1, false>(A, C, st);
    if (dil == 2) return A.dbias ? launch_wgrad3_tma_b<2, 1, true>(A, C, st) : launch_wgrad3_tma_b<2, 1, false>(A, C, st);
  }
  if (C == 8 || C == 16) {
    if (dil == 1) return A.dbias ? launch_wgrad3_tma_b<1, 2, true>(A, C, st) : launch_wgrad3_tma_b<1, 2, false>(A, C, st);
    if (dil == 2) return A.dbias ? launch_wgrad3_tma_b<2, 2, true>(A, C, st) : launch_wgrad3_tma_b<2, 2, false>(A, C, st);
  }
  return NAS3D_ERR_UNSUPPORTED;
}

int tiled_wgrad3_s1(int C, int dil, const WgradArgs& A, cudaStream_t st) {
  if (A.W < 4 || A.ldx % 4 || A.ldy % 4 || !aligned16(A.x) || !aligned16(A.dy)) return NAS3D_ERR_UNSUPPORTED;
  if (C % 4 || C > 64) return NAS3D_ERR_UNSUPPORTED;
  // 32-bit in-sample offsets in the staging loops
  if ((long long)A.Dx * A.Hx * A.Wx * A.ldx >= (1ll << 31) || (long long)A.Dy * A.Hy * A.Wy * A.ldy >= (1ll << 31))
    return NAS3D_ERR_UNSUPPORTED;
  // tile (TH x TD) with the least zero-padded work: 12 x 4 (default), 6 x 8 (e.g. H = 64: 66
  // instead of 72 padded rows; H = 16: 18 instead of 24), 6 x 4 for shallow volumes
  auto padded = [&](int th, int td) {
    return (long long)((A.H + th - 1) / th * th) * ((A.D + td - 1) / td * td);
  };
  {
    const int rc = launch_wgrad3_tma(C, dil, A, st);
    if (rc != NAS3D_ERR_UNSUPPORTED) return rc;
  }
  const long long p12 = padded(12, 4), p68 = padded(6, 8), p64 = padded(6, 4);
  const int cfg = (p12 <= p68 && p12 <= p64) ? 0 : (p68 <= p64 ? 1 : 2);
#define NAS3D_WG(TWT)                                                              \
  if (cfg == 0) {                                                                   \
    if (dil == 1) return launch_wgrad3<1, 12, 4, 16, TWT>(A, C, st);                \
    if (dil == 2) return launch_wgrad3<2, 12, 4, 16, TWT>(A, C, st);                \
  } else if (cfg == 1) {                                                            \
    if (dil == 1) return launch_wgrad3<1, 6, 8, 16, TWT>(A, C, st);                 \
    if (dil == 2) return launch_wgrad3<2, 6, 8, 16, TWT>(A, C, st);                 \
  } else {                                                                          \
    if (dil == 1) return launch_wgrad3<1, 6, 4, 8, TWT>(A, C, st);                  \
    if (dil == 2) return launch_wgrad3<2, 6, 4, 8, TWT>(A, C, st);                  \
  }
  if (A.W <= 8) { NAS3D_WG(8) }
  else if (A.W <= 16) { NAS3D_WG(16) }
  else { NAS3D_WG(32) }
#undef NAS3D_WG
  return NAS3D_ERR_UNSUPPORTED;
}

}  // namespace nas3d
