// argument blocks of the tiled 3x3x3 kernels (conv_tiled.cu), shared with conv_direct.cu
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../include/nas3d_b200.h"

namespace nas3d {

// The kernels work on a LATTICE of D x H x W points.  Tensor x holds lattice point (d,h,w) at
// voxel (d*xs, h*xs, w*xs) of a (Dx,Hx,Wx) volume, tensor y at (d*ys,...) of (Dy,Hy,Wy):
// xs = ys = 1 is the plain stride-1 conv; a stride-2 dilation-2 conv touches only the even
// voxels of its big tensor, i.e. it IS a stride-1 dilation-1 conv between the
// small tensor and the even sub-lattice of the big one (xs or ys = 2).
struct TiledArgs {
  const float* x;      // input (fwd: x, dgrad: dy)
  const float* w;      // W[C][C][27]
  const float* bias;   // fwd only
  float* y;            // output (fwd: y, dgrad: dx)
  int N, D, H, W;      // lattice extents
  int ldx, ldy;
  int accumulate;
  int xs, Dx, Hx, Wx;
  int ys, Dy, Hy, Wy;
  int tiles_w, tiles_h, tiles_d;
  double* moments;     // optional [N][C][2] {sum, sum^2} of the written output (fused GN stats)
  int tma_merged;      // set by the launcher: 4-D tensor map with the (W, C) dims merged (dense C = 4)
};

struct WgradArgs {
  const float* x;
  const float* dy;
  float* dW;
  float* dbias;
  int N, D, H, W;      // lattice extents (= extents of the dy side)
  int ldx, ldy;
  int xs, Dx, Hx, Wx;  // x lives on a lattice of stride xs inside a (Dx,Hx,Wx) volume
  int ys, Dy, Hy, Wy;
  int tiles_w, tiles_h, tiles_d;
};


// stride-2 dilation-1 family (conv_tiled_s2.cu): conv view big_pos = 2*small_pos - 1 + tap
struct S2Args {
  float* big;
  float* small;
  const float* w;
  const float* bias;
  float* dW;
  float* dbias_small;
  int N;
  int Db, Hb, Wb, Cb, ld_big;
  int Ds, Hs, Ws, Cs, ld_small;
  int accumulate;
  int tiles_w, tiles_h, tiles_d;
  double* moments;     // optional fused GN statistics of the produced tensor
};
int tiled_s2_sfb(const S2Args& A, cudaStream_t st);
int tiled_s2_bfs(const S2Args& A, cudaStream_t st);
int tiled_s2_wgrad(const S2Args& A, cudaStream_t st);

// 1x1x1 convs (conv_pointwise.cu); cat != NULL: the big tensor is a virtual concat of dense parts
struct PwCat {
  int nparts;
  const float* const* src;     // parts read as the big tensor (fwd, wgrad)
  float* const* dst;           // parts written as the big gradient (dgrad)
  const float* const* mask;    // relu-mask parts (dgrad), may be NULL
  const int* ld;               // pitch of src / dst parts
  const int* mask_ld;
  const int* acc;              // per-part accumulate flags (dgrad)
};
int pointwise_sfb(const nas3d_conv_desc* d, const float* big, const float* w, const float* bias,
                  const float* scale, int relu, int sigmoid, float* small, int accumulate,
                  double* moments, cudaStream_t st, const PwCat* cat = nullptr);
// opt-in ring-staged forward (conv_pointwise_fwd_ring.cu); NAS3D_ERR_UNSUPPORTED = not taken
int pointwise_sfb_ring(const nas3d_conv_desc* d, const float* big, const float* w, const float* bias,
                       const float* scale, int relu, int sigmoid, float* small, double* moments,
                       cudaStream_t st, const PwCat* cat);
int pointwise_bfs(const nas3d_conv_desc* d, const float* small, const float* w, const float* bias,
                  const float* mask_big, int ld_mask, const float* scale, float* big,
                  int accumulate, cudaStream_t st, const PwCat* cat = nullptr);
int pointwise_wgrad(const nas3d_conv_desc* d, const float* small, const float* big,
                    const float* scale, int relu, float* dW, float* dbias_small, cudaStream_t st,
                    const PwCat* cat = nullptr);

// depthwise 3x3x3 (conv_tiled_dw.cu); wgrad uses S2Args also for stride 1 (big == small extents)
int tiled_dw_s1(bool flip, const TiledArgs& A, int C, cudaStream_t st);
int tiled_dw_s2_sfb(const S2Args& A, cudaStream_t st);
int tiled_dw_s2_bfs(const S2Args& A, cudaStream_t st);
int tiled_dw_wgrad(const S2Args& A, int stride, cudaStream_t st);
int tiled_dw_s2_wgrad_tma(const S2Args& A, cudaStream_t st);   // conv_tiled_s2.cu (TMA tile ring)

// tcgen05 weight gradient of the wide dense 3x3x3 convs (conv_umma_wgrad.cu)
int umma_wgrad(const nas3d_conv_desc* d, const float* small, const float* big, float* dW,
               float* workspace, long long ws_floats, cudaStream_t st);
long long umma_wgrad_workspace_floats(const nas3d_conv_desc* d);

// TMA tensor maps of NDHWC fp32 tensors (conv_tiled.cu); false = not available (option tma = 0,
// no driver entry point, or the box does not fit)
bool make_ndhwc_map(CUtensorMap* m, const float* base, int C, int W, int H, int D, int N, int ld,
                    int bw, int bh, int bd, int bc = 4);
bool make_ndhwc_merged_map(CUtensorMap* m, const float* base, int c, int W, int H, int D, int N,
                           int bw, int bh, int bd);

// return NAS3D_ERR_UNSUPPORTED (without error text) when the shape is not covered
int tiled_conv3_s1(int C, int dil, bool flip, const TiledArgs& A, cudaStream_t st);
int tiled_wgrad3_s1(int C, int dil, const WgradArgs& A, cudaStream_t st);
int fill_channels(float* y, const float* bias, long long nvox, int C, int ld, cudaStream_t st);

}  // namespace nas3d
