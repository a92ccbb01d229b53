/*
 * nas3d_b200.h — C-ABI of the B200-native compute path for nas_3d_unet.
 *
 * The reference (woodywff/nas_3d_unet) is pure PyTorch: its "FFI" for this path is the
 * ATen dispatch underneath torch.nn (prim_ops.py:58,63,66,95-110,133-147,161-163;
 * cell.py:30,32,81,82; loss.py:13-14).  Every entry point below replaces one of those
 * dispatch sites (cited per function) with a hand-written sm_100a kernel.
 *
 * Conventions
 *  - plain pointers and sizes only; all pointers are DEVICE pointers unless marked "host".
 *  - activations are fp32 NDHWC ("channels-last 3-D"): element (n,d,h,w,c) of a tensor
 *    with voxel pitch `ld` (floats) lives at  base[(((n*D+d)*H+h)*W+w)*ld + c].
 *    `ld >= C` lets a tensor be a channel slice of a wider buffer (zero-copy concat,
 *    cell.py:82).  C and ld are multiples of 4 except where stated.
 *  - parameters keep the reference's state_dict layouts: Conv3d [Cout,Cin/g,k,k,k],
 *    ConvTranspose3d [Cin,Cout/g,k,k,k], GroupNorm/Linear as in torch.
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*), allocates
 *    nothing, keeps no pointer after returning, and may be called from any host thread
 *    (the autograd worker thread calls the *_bwd entry points).
 *  - return value: 0 on success, negative nas3d_status on error; nas3d_last_error()
 *    returns a thread-local message.
 */
#ifndef NAS3D_B200_H
#define NAS3D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  NAS3D_OK = 0,
  NAS3D_ERR_ARG = -1,      /* bad shape / alignment / unsupported combination */
  NAS3D_ERR_CUDA = -2,     /* a CUDA runtime call or launch failed */
  NAS3D_ERR_UNSUPPORTED = -3
} nas3d_status;

int nas3d_version(void);
const char* nas3d_last_error(void);
/* number of kernels this library has launched in this process (bench.py: gpu_launches) */
unsigned long long nas3d_launch_count(void);
/* launches of ONE kernel variant, by the label the dispatcher counts it under (e.g.
 * "conv3_s1_tma_merged", "wgrad3_s1", "affine_sum_bwd_apply_ring", "umma_conv"): lets the parity
 * tests assert WHICH kernel served a call.  nas3d_launch_labels writes the newline-separated list of
 * labels seen so far into buf (cap bytes, NUL-terminated) and returns the bytes needed. */
unsigned long long nas3d_launch_count_of(const char* label);
int nas3d_launch_labels(char* buf, int cap);

/* Kernel-selection options (csrc/options.cu).  The reference has one code path per op; here
 * several kernels may serve the same op (tiled vs generic gather, cp.async-ring vs register
 * staging, TMA vs cp.async halo tiles) and every variant computes the same result - the parity
 * tests flip these and compare.  The table is process-wide, read once from the environment
 * (NAS3D_<UPPERCASE NAME>) when the library is loaded and changed afterwards only here; no launch
 * path reads the environment.  Names: tiled, tma, tma_merged, affine_ring, apply_ring,
 * reduce_ring, pw_fwd_ring, reduce_waves, ring_min_log2, pw_vpt_sfb, pw_vpt_bfs, pw_vpt_mom,
 * umma_split_k, umma_wgrad, umma_wgrad_min_c, umma_ws, s1_wgrad_tma,
 * s2_wgrad_tma.  set: 0 or NAS3D_ERR_ARG
 * (unknown name / value out of range); get: the value (>= 0) or NAS3D_ERR_ARG. */
/* Roofline probe: one launch of pure packed-fp32-FMA chains (no memory traffic) over all SMs;
 * returns the flops it executes (> 0) or a negative status.  bench.py times it with CUDA events
 * to state the fp32-FMA roof of THIS run next to the FFMA-bound conv kernels.  out: scratch of at
 * least 148*8*256 floats. */
long long nas3d_probe_fma(float* out, long long out_floats, int iters, void* stream);

int nas3d_set_option(const char* name, int value);
int nas3d_get_option(const char* name);

/* ---------------------------------------------------------------------------------------
 * Layout: NCDHW (what search.py:212 hands over) -> NDHWC with pitch ld_dst.
 * V = D*H*W.  C arbitrary.
 * ------------------------------------------------------------------------------------- */
int nas3d_ncdhw_to_ndhwc(const float* src, float* dst, int N, int C, long long V, int ld_dst,
                         void* stream);

/* ---------------------------------------------------------------------------------------
 * Convolutions.  Replaces nn.Conv3d / nn.ConvTranspose3d forward and convolution_backward
 * (prim_ops.py:95-110,142-147).
 *
 * One descriptor describes the CONV VIEW of the op: a "big" tensor (Db,Hb,Wb,Cb) that is
 * gathered from and a "small" tensor (Ds,Hs,Ws,Cs) with
 *       big_pos = small_pos*stride - pad + tap*dil            (per axis, tap in [0,k))
 *   Conv3d:           big = module input,  small = module output
 *   ConvTranspose3d:  big = module output, small = module input
 * and a weight tensor W[Cs][Cb/groups][k][k][k] (exactly the torch layout in both cases).
 * groups is 1 or Cb==Cs (depthwise).
 * ------------------------------------------------------------------------------------- */
typedef struct {
  int N;
  int Db, Hb, Wb, Cb, ld_big;
  int Ds, Hs, Ws, Cs, ld_small;
  int k, stride, dil, pad;
  int depthwise;            /* 0: groups=1, 1: groups=C */
} nas3d_conv_desc;

/* small[o,cs] = bias[cs] + sum_{tap,cb} f(big[o*s-p+t*d, cb]) * W[cs][cb][tap]
 * f = optional prologue on the big tensor: relu (preprocess 'act_weight_norm', cell.py:47-50)
 * and/or per-(n,cb) scale `big_scale` [N,Cb] (Dropout3d mask, nas.py:50; may be NULL).
 * out_sigmoid: apply 1/(1+exp(-v)) in the epilogue (nn.Sigmoid of last_conv, nas.py:52).
 * Used for: Conv3d forward, ConvTranspose3d dgrad.
 * moments (optional, [N][Cs][2] fp64): receives {sum, sum of squares} per (n,c) of the tensor
 * just written (= nas3d_moments_nc of it; GroupNorm statistics fused into the conv epilogue
 * where the kernel supports it, a separate pass otherwise).  Zeroed inside the call. */
int nas3d_conv_small_from_big(const nas3d_conv_desc* d, const float* big, const float* w,
                              const float* bias, const float* big_scale, int big_relu,
                              int out_sigmoid, float* small, int accumulate, double* moments,
                              void* stream);

/* big[i,cb] (+)= bias[cb] + sum over (o,tap) with o*s-p+t*d == i of small[o,cs]*W[cs][cb][tap]
 * Optional epilogue for the dgrad of a prologue'd Conv3d: result *= (mask_big>0) and
 * *= big_scale[n,cb].  Used for: ConvTranspose3d forward, Conv3d dgrad. */
int nas3d_conv_big_from_small(const nas3d_conv_desc* d, const float* small, const float* w,
                              const float* bias, const float* mask_big, int ld_mask,
                              const float* big_scale, float* big, int accumulate, double* moments,
                              void* stream);

/* dW[cs][cb][tap] += sum_{n,o} small[o,cs] * f(big[o*s-p+t*d, cb]);
 * d_bias_small[cs] += sum small (if non-NULL);  d_bias_big[cb] += sum big (if non-NULL).
 * dW / d_bias must be zero-initialised by the caller (they are accumulated with atomics so
 * they can live directly in the flat gradient bucket that NCCL all-reduces).
 * _ws variant: with `workspace` (workspace_floats >= nas3d_conv_wgrad_workspace_floats(d, prologue
 * present); that query returns 0 for shapes that need none) the wide dense 3x3x3 convs (C = 16 /
 * 32 / 64) run as a tcgen05 split-K GEMM whose K splits are reduced in a fixed order through the
 * workspace (deterministic, no atomics on dW).  The workspace is scratch: owned by the caller
 * (torch's allocator), not retained. nas3d_conv_wgrad == _ws without a workspace. */
long long nas3d_conv_wgrad_workspace_floats(const nas3d_conv_desc* d, int has_prologue);
int nas3d_conv_wgrad_ws(const nas3d_conv_desc* d, const float* small, const float* big,
                        const float* big_scale, int big_relu, float* dW, float* d_bias_small,
                        float* d_bias_big, float* workspace, long long workspace_floats, void* stream);
int nas3d_conv_wgrad(const nas3d_conv_desc* d, const float* small, const float* big,
                     const float* big_scale, int big_relu, float* dW, float* d_bias_small,
                     float* d_bias_big, void* stream);

/* ---------------------------------------------------------------------------------------
 * 1x1x1 convolutions over a VIRTUAL CONCAT: the big tensor is `nparts` (<= 4) dense parts of
 * Cb/nparts channels each (the node outputs of a cell, cell.py:82) that are never copied into
 * one buffer.  Consumers of a cell output are always 1x1x1 convs (cell.py:47-50, nas.py:50).
 * Host arrays of length nparts: part pointers, pitches, per-part accumulate flags.
 * ------------------------------------------------------------------------------------- */
int nas3d_conv1x1_cat_fwd(const nas3d_conv_desc* d, int nparts, const float* const* big_parts,
                          const int* part_ld, const float* w, const float* bias,
                          const float* big_scale, int big_relu, int out_sigmoid, float* small,
                          double* moments, void* stream);
int nas3d_conv1x1_cat_dgrad(const nas3d_conv_desc* d, int nparts, float* const* dbig_parts,
                            const int* part_ld, const int* accumulate, const float* small,
                            const float* w, const float* const* mask_parts, const int* mask_ld,
                            const float* big_scale, void* stream);
int nas3d_conv1x1_cat_wgrad(const nas3d_conv_desc* d, int nparts, const float* const* big_parts,
                            const int* part_ld, const float* small, const float* big_scale,
                            int big_relu, float* dW, float* d_bias_small, void* stream);

/* Fused backward of a dense 1x1x1 STRIDE-1 convolution (the cell preprocess convs cell.py:47-50,
 * separable pointwise convs prim_ops.py:97-98, the head nas.py:50-52): in one pass over memory
 *   dbig[v,cb] (+)= big_scale[n,cb] * (big[v,cb] > 0 if big_relu) * sum_cs ds[v,cs] * W[cs][cb]
 *   dW[cs][cb]  += sum_v ds[v,cs] * f(big[v,cb])       (f = relu and/or scale, as in the forward)
 *   d_bias_small[cs] += sum_v ds[v,cs]                  (if non-NULL)
 * with ds = dsmall, or ds = dsmall * prob * (1 - prob) when small_prob (the sigmoid output of
 * this conv, same layout as dsmall) is given - the backward of nn.Sigmoid (nas.py:52) folded in.
 * It replaces the pair nas3d_conv1x1_cat_dgrad / nas3d_conv_big_from_small + nas3d_conv1x1_cat_wgrad
 * / nas3d_conv_wgrad (+ nas3d_sigmoid_bwd), which read dsmall and big twice.  big is nparts
 * (1..4) dense parts of Cb/nparts channels (nparts = 1: an ordinary tensor with pitch part_ld[0]);
 * dbig_parts NULL (or dbig_parts[0] NULL) = weight/bias gradients only.  Covered: Cs <= 8,
 * Cb <= 64 with Cb/4 dividing 192; nas3d_conv1x1_bwd_fused_supported() tells (1/0). */
int nas3d_conv1x1_bwd_fused_supported(const nas3d_conv_desc* d, int nparts);
int nas3d_conv1x1_bwd_fused(const nas3d_conv_desc* d, int nparts, const float* const* big_parts,
                            const int* part_ld, float* const* dbig_parts, const int* dpart_ld,
                            const int* accumulate, const float* dsmall, const float* small_prob,
                            const float* w, const float* big_scale, int big_relu, float* dW,
                            float* d_bias_small, void* stream);

/* ---------------------------------------------------------------------------------------
 * Tensor-core path (tcgen05.mma kind::tf32, fp32 accumulators in TMEM, 3xTF32 error
 * compensation => fp32-grade results) for dense 3x3x3 convolutions with Cb == Cs in {16,32,64},
 * any stride / dilation, both gather directions (same conv view as above).
 *   nas3d_umma_packed_floats : size (floats) of the packed [W_hi | W_lo] operand, 0 = shape not
 *                              covered (the caller then uses the entry points above)
 *   nas3d_umma_pack_weights  : W[Cs][Cb][27] -> packed K-major core-matrix layout; produce_big
 *                              selects which channel dim is reduced (0: Cb, 1: Cs)
 *   nas3d_umma_conv          : produce_big = 0: small = bias + conv(big)   (Conv3d fwd, ConvT dgrad)
 *                              produce_big = 1: big (+)= bias + convT(small) (ConvT fwd, Conv3d dgrad)
 * ------------------------------------------------------------------------------------- */
long long nas3d_umma_packed_floats(const nas3d_conv_desc* d, int produce_big);
/* packing order of the weight operand: 0 = produce_big 0; 1 = produce_big 1 in tap order; 2 =
 * produce_big 1 in parity-class order (stride-2 transposed gathers with even big extents walk
 * only the (1+pd)(1+ph)(1+pw) taps that exist for each of the 8 output parity classes). */
int nas3d_umma_pack_mode(const nas3d_conv_desc* d, int produce_big);
/* all weight operands of one forward/backward in one launch per 48 entries: host arrays of n
 * channel counts (16/32/64, Cb == Cs), packing modes (above), W pointers and destinations
 * (nas3d_umma_packed_floats floats each, 16-byte aligned).  Weights only change in the
 * optimiser step (search.py:238, train.py:127), so a step needs exactly one such call. */
int nas3d_umma_pack_weights_batch(int n, const int* channels, const int* modes,
                                  const float* const* w, float* const* packed, void* stream);
int nas3d_umma_pack_weights(const nas3d_conv_desc* d, const float* w, int produce_big,
                            float* packed, void* stream);
int nas3d_umma_conv(const nas3d_conv_desc* d, int produce_big, const float* src,
                    const float* packed_w, const float* bias, float* dst, int accumulate,
                    double* moments, void* stream);

/* ---------------------------------------------------------------------------------------
 * Per-(n,c) moments: S[n][c] = {sum x, sum x^2} in fp64 (zeroed inside the call).
 * Serves GroupNorm statistics (prim_ops.py:56-58,77) and the SE squeeze
 * (AdaptiveAvgPool3d, prim_ops.py:133,150).
 * ------------------------------------------------------------------------------------- */
int nas3d_moments_nc(const float* x, int N, long long V, int C, int ld, double* S, void* stream);

/* GroupNorm coefficients: y = a[n,c]*x + b[n,c]  with a = rstd*gamma, b = beta - mean*a.
 * G groups of C/G channels, biased variance, eps (nn.GroupNorm).  mean_rstd [N,G,2] fp32 out. */
int nas3d_gn_coef(const double* S, const float* gamma, const float* beta, int N, int C, int G,
                  long long V, float eps, float* a, float* b, float* mean_rstd, void* stream);

/* SE excitation (prim_ops.py:134-139,151): s[n,c] = sigmoid(W2[c]*relu(W1.mean[n,:]+b1)+b2[c]).
 * W1 [1,C], b1 [1], W2 [C,1], b2 [C].  hz [N,2] = {relu(z), z} saved for backward. */
int nas3d_se_excite(const double* S, const float* W1, const float* b1, const float* W2,
                    const float* b2, int N, int C, long long V, float* s, float* hz, void* stream);

/* ---------------------------------------------------------------------------------------
 * Fused affine-sum:  out = sum_k w_k * act_k(a_k[n,c]*x_k + b_k[n,c])
 * This one kernel is MixedOp's softmax(alpha)-weighted sum (cell.py:30,32) fused with the
 * GroupNorm-apply / ReLU / SE-scale epilogues of the candidate ops (prim_ops.py:75-80,152),
 * the node sum (cell.py:81, searched.py:50) and the concat (cell.py:82: `out` may be a
 * channel slice).  All arrays below are HOST arrays of length nterms (<= NAS3D_MAX_TERMS).
 * a[k]/b[k] may be NULL (=1 / =0); w[k] is a DEVICE pointer to one float or NULL (=1).
 * ------------------------------------------------------------------------------------- */
#define NAS3D_MAX_TERMS 32
int nas3d_affine_sum_fwd(int nterms, const float* const* x, const int* ld_x,
                         const float* const* a, const float* const* b, const float* const* w,
                         const int* relu, float* out, int ld_out, int N, long long V, int C,
                         void* stream);

/* Backward pass 1: per term k and (n,c):  R[k][n][c] = { sum m*dout, sum m*dout*x_k } (fp64),
 * m = relu mask (a*x+b > 0) or 1.  R[k] are device pointers to [N,C,2] doubles (zeroed inside). */
int nas3d_affine_sum_bwd_reduce(int nterms, const float* const* x, const int* ld_x,
                                const float* const* a, const float* const* b, const int* relu,
                                const float* dout, int ld_dout, double* const* R, int N,
                                long long V, int C, void* stream);

/* GroupNorm backward coefficients from R (one term):
 *   dx = p*m*dout + q*x + r ;  dgamma[c] += .. ; dbeta[c] += .. ; dw += <dout, y>  (if dw!=NULL)
 * w: device scalar weight of the term (NULL = 1).  p,q,r: [N,C] fp32 out.
 * dbias_prev (optional, with the forward moments S of x): += the bias gradient of the convolution
 * that produced x, obtained analytically as sum_v dx = p*R1 + q*sum(x) + r*V. */
int nas3d_gn_bwd_coef(const double* R, const float* mean_rstd, const float* gamma,
                      const float* a, const float* b, const float* w, int N, int C, int G,
                      long long V, float* p, float* q, float* r, float* dgamma, float* dbeta,
                      float* dw, const double* S, float* dbias_prev, void* stream);

/* Batched forms of nas3d_gn_coef / nas3d_gn_bwd_coef: all GroupNorm terms of one node (same N, C,
 * G, V) in one launch.  Arrays are HOST arrays of nterms device pointers. */
int nas3d_gn_coef_batch(int nterms, const double* const* S, const float* const* gamma,
                        const float* const* beta, int N, int C, int G, long long V, float eps,
                        float* const* a, float* const* b, float* const* mean_rstd, void* stream);
int nas3d_gn_bwd_coef_batch(int nterms, const double* const* R, const float* const* mean_rstd,
                            const float* const* gamma, const float* const* a,
                            const float* const* b, const float* const* w, int N, int C, int G,
                            long long V, float* const* p, float* const* q, float* const* r,
                            float* const* dgamma, float* const* dbeta, float* const* dw,
                            const double* const* S, float* const* dbias_prev, void* stream);

/* SE backward coefficients from R (term x*s): p = w*s, q = 0, r = dmean/V; parameter grads of
 * the two Linear layers are accumulated (atomics).  S = forward moments of x. */
int nas3d_se_bwd_coef(const double* R, const double* S, const float* s, const float* hz,
                      const float* W1, const float* W2, const float* w, int N, int C, long long V,
                      float* p, float* r, float* dW1, float* db1, float* dW2, float* db2,
                      float* dw, void* stream);

/* Plain term (a=1,b=0, e.g. pooling outputs): only dw += sum_{n,c} R2 is needed. */
int nas3d_plain_bwd_coef(const double* R, const float* w, int N, int C, float* dw, void* stream);

/* Backward pass 2: dx_k (+)= p_k*m*dout + q_k*x_k + r_k.  p/q/r may be NULL (p: use w_k or 1;
 * q,r: 0).  accumulate[k] != 0 adds into dx_k. Terms that share a dx buffer are handled in
 * order by the same thread. */
int nas3d_affine_sum_bwd_apply(int nterms, const float* const* x, const int* ld_x,
                               const float* const* a, const float* const* b, const int* relu,
                               const float* const* p, const float* const* q,
                               const float* const* r, const float* const* w,
                               float* const* dx, const int* ld_dx, const int* accumulate,
                               const float* dout, int ld_dout, int N, long long V, int C,
                               void* stream);

/* The two passes above with the GroupNorm coefficient kernels of their terms FOLDED IN (one launch
 * instead of two on the critical path of every cell node; the coefficient math is a few hundred
 * flops per sample, done by every CTA in its prologue):
 *   nas3d_affine_sum_fwd_gn: for a term k with gn_S[k] != NULL, a[k] / b[k] / gn_mean_rstd[k] are
 *     OUTPUTS computed as by nas3d_gn_coef from the moments gn_S[k] ([N,C,2] fp64), gn_gamma[k],
 *     gn_beta[k] (G groups, eps) and then used; other terms behave as in nas3d_affine_sum_fwd.
 *   nas3d_affine_sum_bwd_apply_gn: for a term k with gn_R[k] != NULL, p[k] / q[k] / r[k] ([N,C]
 *     fp32 scratch) are computed as by nas3d_gn_bwd_coef from gn_R[k] (the output of
 *     nas3d_affine_sum_bwd_reduce), gn_mean_rstd[k], gn_gamma[k], a[k], b[k], w[k], and
 *     gn_dgamma[k] / gn_dbeta[k] / gn_dw[k] (d alpha, may be NULL) / gn_dbias_prev[k] (with
 *     gn_S[k]; may be NULL) are accumulated exactly once. */
int nas3d_affine_sum_fwd_gn(int nterms, const float* const* x, const int* ld_x,
                            float* const* a, float* const* b, const float* const* w,
                            const int* relu, const double* const* gn_S,
                            const float* const* gn_gamma, const float* const* gn_beta,
                            float* const* gn_mean_rstd, int G, float eps, float* out, int ld_out,
                            int N, long long V, int C, void* stream);
int nas3d_affine_sum_bwd_apply_gn(int nterms, const float* const* x, const int* ld_x,
                                  const float* const* a, const float* const* b, const int* relu,
                                  float* const* p, float* const* q, float* const* r,
                                  const float* const* w, float* const* dx, const int* ld_dx,
                                  const int* accumulate, const float* dout, int ld_dout,
                                  const double* const* gn_R, const float* const* gn_mean_rstd,
                                  const float* const* gn_gamma, float* const* gn_dgamma,
                                  float* const* gn_dbeta, float* const* gn_dw,
                                  const double* const* gn_S, float* const* gn_dbias_prev, int G,
                                  int N, long long V, int C, void* stream);

/* ---------------------------------------------------------------------------------------
 * 2x2x2 stride-2 pooling (prim_ops.py:160-168).  kind 0 = avg, 1 = max.
 * Backward of max re-derives the arg-max from x (first maximum in d,h,w scan order, as
 * torch's max_pool3d_with_indices does).
 * ------------------------------------------------------------------------------------- */
int nas3d_pool2_fwd(int kind, const float* x, int ld_x, float* y, int ld_y, int N, int Do, int Ho,
                    int Wo, int C, void* stream);
int nas3d_pool2_bwd(int kind, const float* x, int ld_x, const float* dy, int ld_dy, float* dx,
                    int ld_dx, int accumulate, int N, int Do, int Ho, int Wo, int C, void* stream);

/* ---------------------------------------------------------------------------------------
 * Soft Dice (loss.py:12-14).  pred/truth are addressed as base[n*sn + c*sc + v*sv] so both
 * NDHWC model outputs and NCDHW label tensors are read in place.
 * sums [N,C,3] fp64 = {sum p*t, sum p, sum t} (zeroed inside); loss: 1 float.
 * ------------------------------------------------------------------------------------- */
/* truth: fp32, or int8 {0,1} masks when truth_is_int8 (what generator.py:230-248 produces; a
 * quarter of the host->device bytes of an fp32 upload) */
int nas3d_dice_fwd(const float* pred, long long p_sn, long long p_sc, long long p_sv,
                   const void* truth, int truth_is_int8, long long t_sn, long long t_sc,
                   long long t_sv, int N, int C, long long V, float smooth, double* sums, float* loss,
                   void* stream);
/* dpred[n,c,v] = gout * d loss / d pred, written with pred's addressing. */
int nas3d_dice_bwd(const double* sums, const float* gout, const void* truth, int truth_is_int8,
                   long long t_sn, long long t_sc, long long t_sv, float* dpred, long long p_sn,
                   long long p_sc, long long p_sv, int N, int C, long long V, float smooth,
                   void* stream);

/* dlogit = dprob * prob * (1-prob)   (backward of nn.Sigmoid, nas.py:52); n contiguous floats */
int nas3d_sigmoid_bwd(const float* prob, const float* dprob, float* dlogit, long long n,
                      void* stream);

/* y (+)= x, n contiguous floats (gradient accumulation of multiply-used activations) */
int nas3d_add_inplace(float* y, const float* x, long long n, void* stream);

/* ---------------------------------------------------------------------------------------
 * Sliding-window whole-volume inference on the device (prediction.py:121-170, patches.py:99-206)
 * and label assembly (generator.py:230-248).  Integer results and the float64 stitch are
 * bit-exact with the reference's numpy code (same accumulation order).
 *   extract_patches: volume [C][D][H][W] fp32 + corners [B][3] int32 (device) -> NDHWC batch
 *                    [B][Pd][Ph][Pw][ld_out], zeros outside the volume
 *   patch_nonzero  : flags[b] = 1 iff patch b holds a non-zero voxel (the `np.all(data==0)` test of
 *                    prediction.py:133, on the device: no host sync in the patch loop)
 *   stitch_labels  : patch_preds_dev = DEVICE table of B pointers, patch b = [Pd][Ph][Pw][ld_pred]
 *                    (3 channels; patches may live in different buffers - per forward batch, or
 *                    per rank after an all-gather) -> uint8 labels [D][H][W] (mean of overlaps in
 *                    fp64 in patch order, >= threshold, {1,2,4} assembly, * skull mask, pasted at
 *                    (off_d,off_h,off_w)); patch_nonzero_dev (or NULL): a patch with flag 0 counts
 *                    as an all-zero prediction (prediction.py:134); optional fp64 stitched
 *                    [3][Db][Hb][Wb]
 *   seg_to_masks   : int16 [N][V] -> fp32 [N][3][V] region masks
 * ------------------------------------------------------------------------------------- */
int nas3d_extract_patches(const float* volume, int C, int D, int H, int W, const int* corners_dev,
                          int B, int Pd, int Ph, int Pw, float* out, int ld_out, void* stream);
int nas3d_patch_nonzero(const float* volume, int C, int D, int H, int W, const int* corners_dev, int B,
                        int Pd, int Ph, int Pw, int* flags, void* stream);
int nas3d_stitch_labels(const float* const* patch_preds_dev, const int* patch_nonzero_dev, int ld_pred,
                        const int* corners_dev, int B, int Pd, int Ph, int Pw, int Db, int Hb, int Wb, int D, int H, int W, int off_d, int off_h,
                        int off_w, float threshold, int inclusive, const unsigned char* skull_mask,
                        unsigned char* labels, double* stitched, void* stream);
int nas3d_seg_to_masks(const short* seg, int N, long long V, int inclusive, float* masks, void* stream);

/* Batch staging on the device (generator.py:195-248, augment.py:105-132): x [N][C][D*H*W] planar fp32
 * -> x_out NDHWC (voxel pitch ld_out) with a per-sample index map, seg int16 [N][D*H*W] -> y_out
 * int8 [N][3][D*H*W] region masks (get_multi_class_labels incl. its logical_or quirk).  index_map:
 * HOST int[N][4] = {base, stride_d, stride_h, stride_w}: output voxel (d,h,w) reads source voxel
 * base + stride_d*d + stride_h*h + stride_w*w (signed; identity = {0, H*W, W, 1}).  N <= 64.
 * seg / y_out may both be NULL. */
int nas3d_stage_patches(const float* x, const short* seg, int N, int C, int D, int H, int W,
                        const int* index_map, int inclusive, float* x_out, int ld_out,
                        signed char* y_out, void* stream);
/* same with the segmentation as int8 (the label values 0, 1, 2, 4 fit): 1 byte per voxel over PCIe */
int nas3d_stage_patches_seg8(const float* x, const signed char* seg, int N, int C, int D, int H, int W,
                        const int* index_map, int inclusive, float* x_out, int ld_out,
                        signed char* y_out, void* stream);

/* ---- flat-buffer Adam (replaces torch.optim.Adam of search.py:103-104,228,237; train.py:58,127) ----
 * param / exp_avg / exp_avg_sq: flat fp32 arenas holding every tensor of one param group at offsets
 * that are multiples of 4 floats.  grad_ptrs: HOST array of ntensors device pointers, one gradient
 * per tensor (NULL = no gradient this step, tensor left untouched); they travel as kernel
 * parameters, so a captured graph keeps them.  chunks_dev: DEVICE array of int4 {tensor index, flat
 * offset, offset inside the tensor, length <= nas3d_adam_chunk_floats()}, sorted by tensor;
 * tensor_first_chunk: HOST int[ntensors+1], first chunk of each tensor.  hyper_dev: DEVICE float[16]
 * = {lr, beta1, beta2, eps, weight_decay, step, -, -, maximize, 1-beta1, 1-beta2}; the call increments step and
 * derives the bias corrections on the device (graph-capturable). */
int nas3d_adam_chunk_floats(void);
int nas3d_adam_flat_step(float* param, float* exp_avg, float* exp_avg_sq,
                         const float* const* grad_ptrs, int ntensors, const int* tensor_first_chunk,
                         const int* chunks_dev, float* hyper_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NAS3D_B200_H */
