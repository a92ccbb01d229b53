// Fused backward of a dense 1x1x1 stride-1 convolution (cell preprocess convs, separable
// pointwise convs, the 12->3 head; prim_ops.py:95-110, cell.py:47-50, nas.py:50-52):
//
//   dbig[v,cb]  (+)= scale[n,cb] * mask(big[v,cb]) * sum_cs dsmall[v,cs] * W[cs][cb]      (dgrad)
//   dW[cs][cb]   += sum_v dsmall[v,cs] * f(big[v,cb]),  f = relu and/or per-(n,cb) scale   (wgrad)
//   dbias[cs]    += sum_v dsmall[v,cs]
//   optionally dsmall = dprob * prob * (1 - prob)   (backward of the head's nn.Sigmoid, nas.py:52)
//
// Separately these are two HBM-bound passes that read the same two tensors (dgrad reads dsmall
// and the relu-mask source = big, wgrad reads dsmall and big): 4 tensor reads + 1 write.  Fused,
// one pass reads dsmall and big once and writes dbig: the algorithmic minimum of the pair.
//
// Mapping ("quad" mapping): the wide (big) side is walked as a flat array of float4 channel
// quads, item i = voxel * Q + quad (Q = Cb/4), one item per thread per step.  A warp's 32
// accesses to big / dbig are 512 contiguous bytes for a dense tensor (a thread-per-voxel mapping
// touches 32 x 16 B at a 4*Cb-byte stride: half-used sectors on every store), the narrow dsmall
// vector of a voxel is read by its Q threads through L1.  The block size is a multiple of Q, so
// a thread's quad is fixed: its W[.][quad] slice and its dW partial sums (CS x 4) live in
// registers for the whole kernel; partial sums are combined per CTA in shared memory and
// leave with one atomicAdd per (cs, cb) per CTA (the grid is one resident wave).
#include "common.cuh"
#include "conv_tiled.h"

namespace nas3d {

struct PwBwdArgs {
  const float* dy;       // dsmall [nvox][lds]  (or dprob when prob != NULL)
  const float* prob;     // optional sigmoid output of this conv, same layout as dy
  const float* w;        // W[Cs][Cb]
  const float* scale;    // optional [N][Cb]
  float* dW;             // [Cs][Cb]
  float* dbias;          // optional [Cs]
  const float* x[4];     // parts of big
  float* dx[4];          // parts of dbig (dx[0] == NULL: no dgrad wanted)
  int ldx[4], lddx[4], acc[4];
  int nseg, seg_w;
  int Cs, Cb, lds;
  unsigned nvox, Vs;
  int relu;
};

template <typename T>
__device__ __forceinline__ T pb_pick4(const T (&a)[4], int i) {
  return i == 0 ? a[0] : (i == 1 ? a[1] : (i == 2 ? a[2] : a[3]));
}

constexpr int PB_T = 192;        // divisible by every supported Q (1,2,3,4,6,8,12,16)
#ifndef NAS3D_PB_S
#define NAS3D_PB_S 4
#endif
#ifndef NAS3D_PB_MINB
#define NAS3D_PB_MINB 4
#endif
// cp.async ring depth (tiles in flight per thread: PB_S - 1) and CTAs per SM of the Cs <= 4
// instantiations.  Measured (profiles/r1f_ab_pb_variants.txt): 4 stages x 4 CTAs 416.7 patches/s,
// 6 x 3 415.1, 8 x 3 414.5 - the shallower ring leaves more of the SM's L1 to the dsmall reads.
constexpr int PB_S = NAS3D_PB_S;
constexpr int PB_MAX_CB = 64;

__device__ __forceinline__ unsigned pb_smem(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void pb_cp16(void* dst, const void* src) {      // streaming: L2 only
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(pb_smem(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void pb_cp16_ca(void* dst, const void* src) {   // shared by Q lanes: via L1
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(pb_smem(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void pb_cp4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(pb_smem(dst)), "l"(src) : "memory");
}

// The loads of a thread's next PB_S-1 items are in flight as cp.async copies into the thread's OWN
// slots of a shared-memory ring (no register is held by a pending load, and since a thread only
// reads what it copied itself, cp.async.wait_group is the only synchronisation): at 3 CTAs x 192
// threads a register-staged version had ~18 KB of reads in flight per SM and ran at 2.4-3.7 TB/s
// (profiles/r1f_s1_*); HBM needs ~35 KB per SM.  With 16 bytes of big per item the kernel is also
// instruction-sensitive: a CTA walks a CONTIGUOUS range of tiles so that every address is a
// running pointer (+= constant), the per-(n,cb) scale is reloaded only at sample boundaries, and
// the scalar dsmall path (the head: Cs = 3, pitch 3) is its own instantiation.
template <int CS, bool DX, bool PROB, bool VEC>
__global__ void __launch_bounds__(PB_T, CS == 4 ? NAS3D_PB_MINB : 2) pw_bwd_fused_kernel(const __grid_constant__ PwBwdArgs A, int tiles_per_cta) {
  extern __shared__ __align__(16) float4 ring[];   // [component][stage][thread]
  __shared__ float sW[CS * PB_MAX_CB];
  __shared__ float sB[CS];
  constexpr int D4 = CS / 4;
  const int Cb = A.Cb, Cs = A.Cs;
  const int Q = Cb >> 2;
  // quad of this thread.  Dense big tensor: consecutive threads walk consecutive quads (a warp =
  // 512 contiguous bytes).  Virtual concat of 4-channel parts: a warp stays inside ONE part and
  // walks 32 consecutive voxels of it, which is what is contiguous there.
  const bool part_major = (A.seg_w == 4) && (A.nseg == Q) && (Q > 1);
  const int g = part_major ? threadIdx.x / (PB_T / Q) : threadIdx.x % Q;
  const int c0 = g * 4;
  const int sg = c0 / A.seg_w, off = c0 - sg * A.seg_w;
  const int xl = pb_pick4(A.ldx, sg);
  const int dxl = pb_pick4(A.lddx, sg);
  const bool do_acc = DX && pb_pick4(A.acc, sg) != 0;
  const bool bias_lane = (g == 0) && (A.dbias != nullptr);
  // ring components: x | dy[D4] | prob[D4] (PROB) | old (when any part accumulates), each
  // [stage][thread]: all components of a stage sit at constant offsets from its x slot
  constexpr int O_DY = PB_S * PB_T, O_P = O_DY + D4 * PB_S * PB_T,
                O_OLD = O_P + (PROB ? D4 * PB_S * PB_T : 0);
  float4* const r0 = ring + threadIdx.x;
  float4* const r_end = r0 + PB_S * PB_T;

  for (int i = threadIdx.x; i < CS * Cb; i += PB_T) sW[i] = 0.f;
  if (threadIdx.x < CS) sB[threadIdx.x] = 0.f;

  // this thread's weight slice W[cs][c0..c0+3] (dgrad) and partial sums (wgrad)
  float4 Wr[DX ? CS : 1];
  if (DX) {
#pragma unroll
    for (int cs = 0; cs < CS; ++cs) {
      Wr[cs] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (cs < Cs) {
        const float* p = A.w + (long long)cs * Cb + c0;
        Wr[cs] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
      }
    }
  }
  float2 acc[CS][2];
  float bs[CS];
#pragma unroll
  for (int cs = 0; cs < CS; ++cs) {
    acc[cs][0] = acc[cs][1] = make_float2(0.f, 0.f);
    bs[cs] = 0.f;
  }

  const unsigned vpt = PB_T / Q;                       // voxels per tile
  const unsigned ntiles = (A.nvox + vpt - 1) / vpt;
  const unsigned tile0 = blockIdx.x * (unsigned)tiles_per_cta;
  const unsigned tile1 = min(tile0 + (unsigned)tiles_per_cta, ntiles);
  const unsigned vend = min(tile1 * vpt, A.nvox);      // this CTA's voxels: [tile0*vpt, vend)
  unsigned v = tile0 * vpt + (part_major ? threadIdx.x % vpt : threadIdx.x / Q);   // consume side
  unsigned vi = v;                                     // issue side (PB_S-1 tiles ahead)
  // running pointers of the issue side
  const float* px = pb_pick4(A.x, sg) + off + (long long)vi * xl;
  const float* pdy = A.dy + (long long)vi * A.lds;
  const float* pp = PROB ? A.prob + (long long)vi * A.lds : nullptr;
  const float* pold = DX ? pb_pick4(A.dx, sg) + off + (long long)vi * dxl : nullptr;
  float* pdx = const_cast<float*>(pold);               // consume side
  const long long sx = (long long)vpt * xl, sdy = (long long)vpt * A.lds, sdx = (long long)vpt * dxl;

  // issue the copies of this thread's next item into the ring stage at `is` (one commit group)
  float4* is = r0;
  auto issue = [&]() {
    if (vi < vend) {
      pb_cp16(is, px);
      if (VEC) {
#pragma unroll
        for (int i = 0; i < D4; ++i) pb_cp16_ca(is + O_DY + i * PB_S * PB_T, pdy + 4 * i);
      } else {
#pragma unroll
        for (int i = 0; i < CS; ++i)
          if (i < Cs) pb_cp4(reinterpret_cast<float*>(is + O_DY + (i >> 2) * PB_S * PB_T) + (i & 3), pdy + i);
      }
      if (PROB) {
        if (VEC) {
#pragma unroll
          for (int i = 0; i < D4; ++i) pb_cp16_ca(is + O_P + i * PB_S * PB_T, pp + 4 * i);
        } else {
#pragma unroll
          for (int i = 0; i < CS; ++i)
            if (i < Cs) pb_cp4(reinterpret_cast<float*>(is + O_P + (i >> 2) * PB_S * PB_T) + (i & 3), pp + i);
        }
      }
      if (do_acc) pb_cp16(is + O_OLD, pold);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    vi += vpt; px += sx; pdy += sdy;
    if (PROB) pp += sdy;
    if (DX) pold += sdx;
    is += PB_T;
    if (is == r_end) is = r0;
  };

#pragma unroll
  for (int s = 0; s < PB_S - 1; ++s) issue();
  // per-(n, cb) scale of the current sample; reloaded when v crosses a sample boundary
  const bool has_scale = A.scale != nullptr;
  float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
  unsigned vb = 0xffffffffu;                           // first voxel of the next sample
  if (has_scale && v < vend) {
    const unsigned n = v / A.Vs;
    sc = ldg4(A.scale + (long long)n * Cb + c0);
    vb = (n + 1) * A.Vs;
  }
  const float4* cs_ = r0;                              // consume-side stage
#pragma unroll 1
  for (unsigned t = tile0; t < tile1; ++t) {
    issue();
    asm volatile("cp.async.wait_group %0;\n" ::"n"(PB_S - 1) : "memory");
    if (v < vend) {
      if (v >= vb) {
        const unsigned n = v / A.Vs;
        sc = ldg4(A.scale + (long long)n * Cb + c0);
        vb = (n + 1) * A.Vs;
      }
      float dyv[CS];
      const float4 xq = cs_[0];
#pragma unroll
      for (int i = 0; i < D4; ++i) {
        const float4 t4 = cs_[O_DY + i * PB_S * PB_T];
        dyv[4 * i] = t4.x; dyv[4 * i + 1] = t4.y; dyv[4 * i + 2] = t4.z; dyv[4 * i + 3] = t4.w;
      }
      if (PROB) {   // dlogit = dprob * p * (1 - p)
#pragma unroll
        for (int i = 0; i < D4; ++i) {
          const float4 p4 = cs_[O_P + i * PB_S * PB_T];
          dyv[4 * i] *= p4.x * (1.f - p4.x); dyv[4 * i + 1] *= p4.y * (1.f - p4.y);
          dyv[4 * i + 2] *= p4.z * (1.f - p4.z); dyv[4 * i + 3] *= p4.w * (1.f - p4.w);
        }
      }
      if (Cs < CS) {   // channels past Cs were never copied / are pad lanes of the pitch
#pragma unroll
        for (int i = 0; i < CS; ++i) dyv[i] = i < Cs ? dyv[i] : 0.f;
      }
      float4 xf = xq;
      if (A.relu) {
        xf.x = fmaxf(xf.x, 0.f); xf.y = fmaxf(xf.y, 0.f);
        xf.z = fmaxf(xf.z, 0.f); xf.w = fmaxf(xf.w, 0.f);
      }
      if (has_scale) { xf.x *= sc.x; xf.y *= sc.y; xf.z *= sc.z; xf.w *= sc.w; }
#pragma unroll
      for (int cs = 0; cs < CS; ++cs) {
        fma2(acc[cs][0], dyv[cs], xf.x, xf.y);
        fma2(acc[cs][1], dyv[cs], xf.z, xf.w);
        if (bias_lane) bs[cs] += dyv[cs];
      }
      if (DX) {
        float2 d0 = make_float2(0.f, 0.f), d1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int cs = 0; cs < CS; ++cs) {
          fma2(d0, dyv[cs], Wr[cs].x, Wr[cs].y);
          fma2(d1, dyv[cs], Wr[cs].z, Wr[cs].w);
        }
        float4 r = make_float4(d0.x, d0.y, d1.x, d1.y);
        if (A.relu) {
          r.x = xq.x > 0.f ? r.x : 0.f; r.y = xq.y > 0.f ? r.y : 0.f;
          r.z = xq.z > 0.f ? r.z : 0.f; r.w = xq.w > 0.f ? r.w : 0.f;
        }
        if (has_scale) { r.x *= sc.x; r.y *= sc.y; r.z *= sc.z; r.w *= sc.w; }
        if (do_acc) {
          const float4 old = cs_[O_OLD];
          r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w;
        }
        st4(pdx, r);
      }
    }
    v += vpt;
    if (DX) pdx += sdx;
    cs_ += PB_T;
    if (cs_ == r_end) cs_ = r0;
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");

  // ---- combine the partial sums of the CTA, then one atomic per (cs, cb) and per cs
  __syncthreads();
#pragma unroll
  for (int cs = 0; cs < CS; ++cs) {
    if (cs < Cs) {
      float* row = sW + cs * Cb + c0;
      atomicAdd(row + 0, acc[cs][0].x);
      atomicAdd(row + 1, acc[cs][0].y);
      atomicAdd(row + 2, acc[cs][1].x);
      atomicAdd(row + 3, acc[cs][1].y);
      if (bias_lane) atomicAdd(&sB[cs], bs[cs]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Cs * Cb; i += PB_T) atomicAdd(A.dW + i, sW[i]);
  if (A.dbias && threadIdx.x < Cs) atomicAdd(A.dbias + threadIdx.x, sB[threadIdx.x]);
}

template <int CS, bool DX, bool PROB, bool VEC>
static int pb_launch(const PwBwdArgs& A, cudaStream_t st) {
  bool any_acc = false;
  for (int i = 0; i < A.nseg; ++i) any_acc = any_acc || (DX && A.acc[i]);
  const int comps = 1 + CS / 4 + (PROB ? CS / 4 : 0) + (any_acc ? 1 : 0);
  const int smem = comps * PB_S * PB_T * (int)sizeof(float4);
  auto kern = pw_bwd_fused_kernel<CS, DX, PROB, VEC>;
  static int occ[2] = {0, 0};   // per instantiation: without / with the `old` ring component
  int& o = occ[any_acc ? 1 : 0];
  if (o == 0) {
    NAS3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    int q = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, kern, PB_T, smem) != cudaSuccess || q < 1) q = 1;
    o = q;
  }
  const unsigned Q = (unsigned)A.Cb / 4;
  const unsigned vpt = PB_T / Q;
  const unsigned ntiles = (A.nvox + vpt - 1) / vpt;
  unsigned gx = (unsigned)o * kNumSMs;     // one resident wave, each CTA a contiguous tile range
  if (gx > ntiles) gx = ntiles;
  if (gx < 1) gx = 1;
  const unsigned tpc = (ntiles + gx - 1) / gx;
  gx = (ntiles + tpc - 1) / tpc;
  kern<<<gx, PB_T, smem, st>>>(A, (int)tpc);
  return NAS3D_OK;
}

static bool pb_supported(const nas3d_conv_desc* d, int nparts) {
  if (!d || d->k != 1 || d->stride != 1 || d->pad != 0 || d->depthwise) return false;
  if (d->Cb % 4 || d->Cb > PB_MAX_CB || d->Cs < 1 || d->Cs > 8) return false;
  if (PB_T % (d->Cb / 4)) return false;
  if (nparts < 1 || nparts > 4 || d->Cb % nparts || (d->Cb / nparts) % 4) return false;
  if (d->Db != d->Ds || d->Hb != d->Hs || d->Wb != d->Ws) return false;
  const long long nvox = (long long)d->N * d->Ds * d->Hs * d->Ws;
  return nvox > 0 && nvox < (1ll << 31) - 4096;
}

}  // namespace nas3d

using namespace nas3d;

extern "C" {

int nas3d_conv1x1_bwd_fused_supported(const nas3d_conv_desc* d, int nparts) {
  return pb_supported(d, nparts) ? 1 : 0;
}

int nas3d_conv1x1_bwd_fused(const nas3d_conv_desc* d, int nparts, const float* const* big_parts,
                            const int* part_ld, float* const* dbig_parts, const int* dpart_ld,
                            const int* accumulate, const float* dsmall, const float* small_prob,
                            const float* w, const float* big_scale, int big_relu, float* dW,
                            float* d_bias_small, void* stream) {
  NAS3D_REQUIRE(pb_supported(d, nparts), "conv1x1_bwd_fused: shape not covered (k=%d s=%d Cb=%d Cs=%d parts=%d)",
                d ? d->k : -1, d ? d->stride : -1, d ? d->Cb : -1, d ? d->Cs : -1, nparts);
  NAS3D_REQUIRE(big_parts && part_ld && dsmall && w && dW, "conv1x1_bwd_fused: NULL argument");
  PwBwdArgs A{};
  A.dy = dsmall; A.prob = small_prob; A.w = w; A.scale = big_scale; A.dW = dW; A.dbias = d_bias_small;
  A.nseg = nparts; A.seg_w = d->Cb / nparts;
  A.Cs = d->Cs; A.Cb = d->Cb; A.lds = d->ld_small;
  A.nvox = (unsigned)((long long)d->N * d->Ds * d->Hs * d->Ws);
  A.Vs = (unsigned)((long long)d->Ds * d->Hs * d->Ws);
  A.relu = big_relu;
  const bool want_dx = dbig_parts != nullptr && dbig_parts[0] != nullptr;
  if (d->ld_small % 4 == 0)
    NAS3D_REQUIRE(aligned16(dsmall) && (!small_prob || aligned16(small_prob)), "conv1x1_bwd_fused: dsmall must be 16-byte aligned");
  NAS3D_REQUIRE(!big_scale || aligned16(big_scale), "conv1x1_bwd_fused: scale must be 16-byte aligned");
  for (int i = 0; i < nparts; ++i) {
    NAS3D_REQUIRE(big_parts[i] && aligned16(big_parts[i]) && part_ld[i] % 4 == 0 && part_ld[i] >= A.seg_w,
                  "conv1x1_bwd_fused: part %d of big: pitch %d / alignment", i, part_ld[i]);
    A.x[i] = big_parts[i]; A.ldx[i] = part_ld[i];
    if (want_dx) {
      NAS3D_REQUIRE(dpart_ld && dbig_parts[i] && aligned16(dbig_parts[i]) && dpart_ld[i] % 4 == 0 &&
                        dpart_ld[i] >= A.seg_w,
                    "conv1x1_bwd_fused: part %d of dbig: pitch / alignment", i);
      A.dx[i] = dbig_parts[i]; A.lddx[i] = dpart_ld[i]; A.acc[i] = accumulate ? accumulate[i] : 0;
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  const bool vec = d->ld_small % 4 == 0 && d->ld_small >= (d->Cs <= 4 ? 4 : 8);   // whole float4s of dsmall
  if (small_prob) {
    NAS3D_REQUIRE(d->Cs <= 4 && want_dx, "conv1x1_bwd_fused: the sigmoid prologue serves the head (Cs <= 4, with dgrad)");
    rc = vec ? pb_launch<4, true, true, true>(A, st) : pb_launch<4, true, true, false>(A, st);
  } else if (d->Cs <= 4) {
    if (vec) rc = want_dx ? pb_launch<4, true, false, true>(A, st) : pb_launch<4, false, false, true>(A, st);
    else rc = want_dx ? pb_launch<4, true, false, false>(A, st) : pb_launch<4, false, false, false>(A, st);
  } else {
    if (vec) rc = want_dx ? pb_launch<8, true, false, true>(A, st) : pb_launch<8, false, false, true>(A, st);
    else rc = want_dx ? pb_launch<8, true, false, false>(A, st) : pb_launch<8, false, false, false>(A, st);
  }
  if (rc) return rc;
  return launched("pw_bwd_fused");
}

}  // extern "C"
