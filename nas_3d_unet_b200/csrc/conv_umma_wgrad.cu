// tcgen05 / TMEM weight gradient of the WIDE dense 3x3x3 convolutions (C = 16, 32, 64; any stride /
// dilation, Conv3d and ConvTranspose3d through the conv view):
//
//     dW[cs][cb][tap] = sum over small voxels o   small[o][cs] * big[o*stride - pad + tap*dil][cb]
//
// as ONE split-K GEMM per launch:   D[m = (tap, cb)][n = cs]  +=  A[m][k = o] * B[n][k = o]
//
// * M = 128 rows = 128/C taps x C big-channels per tile, 27 taps -> 4 / 7 / 14 tiles; N = cs;
//   K = the small voxels of the CTA's chunk.  All tiles of a CTA accumulate in TMEM for the whole
//   chunk (fp32, written ONCE), then the epilogue adds them to dW with atomics (split-K over
//   CTAs; C = 64 additionally splits the taps over gridDim.y because 14 tiles x 128 columns
//   exceed the 512 TMEM columns).
// * Both operands are "MN-major": the NDHWC tensors hold the channel (= M resp. N) index
//   contiguously, so a thread moves 4 channels of one voxel as one 16-byte vector from global
//   memory into the canonical no-swizzle MN-major core-matrix layout [16-B channel group][8 voxels]
//   - no transposition, no scalar stores (instruction descriptor a_major = b_major = 1).
// * 3xTF32 as in conv_umma.cu: x = hi + lo (hi = top 19 bits); B' = [small_hi | small_lo] along N:
//       D[:, 0:2C] += A_hi * B'        D[:, 0:C] += A_lo * B_hi       epilogue: D[:, c] + D[:, C + c]
// * 256 threads gather (voxel = tid % KB fixed per stage, (tap, channel quad) pairs strided over
//   the rest), registers -> smem double buffered, one elected thread issues the MMAs,
//   tcgen05.commit -> mbarrier releases a stage.
// The deep U-Net levels are where this matters: their FFMA wgrad kernels ran at 2-20 TFLOP/s
// (launches of 60-170 us that are pure latency), see DESIGN.md.
#include "common.cuh"
#include "conv_tiled.h"

namespace nas3d {

namespace uw {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "UW_WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra UW_DONE;\n\t"
      "bra UW_WAIT_LOOP;\n\t"
      "UW_DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// SmemDescriptor (sm_100), SWIZZLE_NONE: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version 1.
// MN-major no-swizzle canonical layout (cute mma_traits_sm100.hpp): ((4,1,m),(8,k)):((1,4,SBO),(4,LBO))
// in elements: a 16-byte group of 4 MN-elements is contiguous, the 8 K-rows of a core matrix follow
// at 16-byte pitch, MN groups are SBO apart, 8-row K blocks LBO apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// InstrDescriptor: c_format F32 (1) [4,6) | a_format TF32 (2) [7,10) | b_format TF32 (2) [10,13)
// | a_major MN (1) [15] | b_major MN (1) [16] | n>>3 [17,23) | m>>4 [24,29), M = 128
__device__ __forceinline__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
         ((128u >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                     uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float v[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
                 "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float4 ldg4_pred(const float* p, bool ok) {
  float4 v;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.f32 %0, 0f00000000;\n\t"
      "mov.f32 %1, 0f00000000;\n\t"
      "mov.f32 %2, 0f00000000;\n\t"
      "mov.f32 %3, 0f00000000;\n\t"
      "@p ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n\t"
      "}\n"
      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
      : "l"(p), "r"((int)ok));
  return v;
}
__device__ __forceinline__ void split_hi_lo(const float4 x, float4* hi, float4* lo) {
  hi->x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
  hi->y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
  hi->z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
  hi->w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
  lo->x = x.x - hi->x; lo->y = x.y - hi->y; lo->z = x.z - hi->z; lo->w = x.w - hi->w;
}

}  // namespace uw

struct UwArgs {
  const float* big;
  const float* small;
  float* dW;               // [C][C][27]
  int Db, Hb, Wb, ldb;
  int Ds, Hs, Ws, lds;
  int stride, dil, pad;
  long long nvox;          // N * Ds * Hs * Ws (the K extent of the GEMM)
  int chunk;               // small voxels per CTA, a multiple of KB
  int debug;
};

template <int C>
struct UwShape {
  static constexpr int THREADS = 256;
  static constexpr int KB = C == 32 ? 8 : 16;                // voxels per smem stage
  static constexpr int KBLK = KB / 8;                        // K = 8 MMA steps per stage
  static constexpr int CQ = C / 4;                           // 16-byte channel groups per voxel
  static constexpr int TPT = 128 / C;                        // taps per 128-row tile
  static constexpr int TILES_ALL = (27 + TPT - 1) / TPT;     // 4 / 7 / 14
  static constexpr int TILES = C == 64 ? 4 : TILES_ALL;      // tiles per CTA
  static constexpr int GROUPS = (TILES_ALL + TILES - 1) / TILES;   // gridDim.y: 1 / 1 / 4
  static constexpr int TAPS_CTA = TILES * TPT;               // 32 / 28 / 8 tap slots per CTA
  static constexpr int NCOL = 2 * C;                         // [hi | lo] columns of one tile
  static constexpr int TMEM_COLS = TILES * NCOL <= 128 ? 128 : (TILES * NCOL <= 256 ? 256 : 512);
  static constexpr int A_TILE_BYTES = 128 * KB * 4;          // one tile, one of hi / lo
  static constexpr int A_BYTES = TILES * A_TILE_BYTES;
  static constexpr int B_BYTES = NCOL * KB * 4;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + B_BYTES;
  static constexpr size_t SMEM = 2 * STAGE_BYTES + 64;
  static constexpr int SUBS = THREADS / KB;                  // (tap, quad) pairs walked in parallel
  static constexpr int PAIRS = TAPS_CTA * CQ;                // 16-byte items per voxel per stage
  static constexpr int ITEMS = (PAIRS + SUBS - 1) / SUBS;    // per thread per stage
  static_assert(TILES * NCOL <= 512, "TMEM columns");
  static_assert(SMEM <= 227 * 1024, "shared memory");
};

template <int C>
__global__ void __launch_bounds__(256) umma_wgrad_kernel(const UwArgs A) {
  using US = UwShape<C>;
  constexpr int KB = US::KB, CQ = US::CQ, TILES = US::TILES, NCOL = US::NCOL;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * US::STAGE_BYTES);   // stage drained [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 2 * US::STAGE_BYTES + 32);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    uw::mbar_init(&bars[0], 1);
    uw::mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                     uw::smem_u32(tmem_slot)),
                 "r"((uint32_t)US::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  const int tap0 = (int)blockIdx.y * US::TAPS_CTA;           // first tap slot of this CTA
  const long long k0 = (long long)blockIdx.x * A.chunk;
  long long k1 = k0 + A.chunk;
  if (k1 > A.nvox) k1 = A.nvox;
  const int nstage = (int)((k1 - k0 + KB - 1) / KB);

  const int v = tid % KB, sub = tid / KB;                     // my voxel slot, my first pair
  const uint32_t vrow = (uint32_t)((v >> 3) * 4096 + (v & 7) * 16);   // K block, row inside it (A)
  const uint32_t vrow_b = (uint32_t)((v >> 3) * (NCOL / 4 * 128) + (v & 7) * 16);

  float4 xa[US::ITEMS];
  float4 xb = make_float4(0.f, 0.f, 0.f, 0.f);
  auto gather = [&](int j) {
    const long long o = k0 + (long long)j * KB + v;
    const bool vok = o < k1;
    long long t = vok ? o : 0;
    const int ow = (int)(t % A.Ws); t /= A.Ws;
    const int oh = (int)(t % A.Hs); t /= A.Hs;
    const int od = (int)(t % A.Ds);
    const int n = (int)(t / A.Ds);
    const float* bn = A.big + (long long)n * A.Db * A.Hb * A.Wb * A.ldb;
    const int bd = od * A.stride - A.pad, bh = oh * A.stride - A.pad, bw = ow * A.stride - A.pad;
#pragma unroll
    for (int i = 0; i < US::ITEMS; ++i) {
      const int p = sub + i * US::SUBS;
      const int tap = tap0 + p / CQ, cq = p % CQ;
      const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
      const int id = bd + kd * A.dil, ih = bh + kh * A.dil, iw = bw + kw * A.dil;
      const bool ok = vok && p < US::PAIRS && tap < 27 && (unsigned)id < (unsigned)A.Db &&
                      (unsigned)ih < (unsigned)A.Hb && (unsigned)iw < (unsigned)A.Wb;
      const float* px = ok ? bn + (((long long)id * A.Hb + ih) * A.Wb + iw) * A.ldb + cq * 4 : A.big;
      xa[i] = uw::ldg4_pred(px, ok);
    }
    if (sub < CQ) xb = uw::ldg4_pred(vok ? A.small + o * A.lds + sub * 4 : A.small, vok);
  };

  constexpr uint32_t IDESC_FULL = uw::make_idesc(NCOL);
  constexpr uint32_t IDESC_HALF = uw::make_idesc(C);

  if (nstage > 0) gather(0);
#pragma unroll 1
  for (int j = 0; j < nstage; ++j) {
    const int s = j & 1;
    unsigned char* stage = smem + s * US::STAGE_BYTES;
    unsigned char* a_hi = stage;
    unsigned char* a_lo = stage + US::A_BYTES;
    unsigned char* b_sm = stage + 2 * US::A_BYTES;
    if (j >= 2) uw::mbar_wait(&bars[s], (uint32_t)(((j >> 1) - 1) & 1));   // MMAs of stage j-2 drained
    // A: row m = (tap slot, cb) of tile p / 32: 16-byte group (p % 32), K row v
#pragma unroll
    for (int i = 0; i < US::ITEMS; ++i) {
      const int p = sub + i * US::SUBS;
      if (p < US::PAIRS) {
        float4 hi, lo;
        uw::split_hi_lo(xa[i], &hi, &lo);
        const uint32_t off = (uint32_t)(p >> 5) * US::A_TILE_BYTES + (uint32_t)(p & 31) * 128u + vrow;
        *reinterpret_cast<float4*>(a_hi + off) = hi;
        *reinterpret_cast<float4*>(a_lo + off) = lo;
      }
    }
    // B: columns [0, C) = hi, [C, 2C) = lo; 16-byte column group (cs / 4), K row v
    if (sub < CQ) {
      float4 hi, lo;
      uw::split_hi_lo(xb, &hi, &lo);
      *reinterpret_cast<float4*>(b_sm + (uint32_t)sub * 128u + vrow_b) = hi;
      *reinterpret_cast<float4*>(b_sm + (uint32_t)(CQ + sub) * 128u + vrow_b) = lo;
    }
    if (j + 1 < nstage) gather(j + 1);          // in flight while this stage's MMAs are issued
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      if (A.debug && blockIdx.x == 0 && blockIdx.y == 0 && j == 0) {
        const float* fa = reinterpret_cast<const float*>(a_hi);
        const float* fb = reinterpret_cast<const float*>(b_sm);
        printf("uw dbg: tmem_base %08x nstage %d chunk %d a_hi[0..3] %g %g %g %g  a_hi[row1] %g  b[0..3] %g %g %g %g\n",
               tmem_base, nstage, A.chunk, fa[0], fa[1], fa[2], fa[3], fa[4], fb[0], fb[1], fb[2], fb[3]);
      }
      const uint32_t a_hi_s = uw::smem_u32(a_hi), a_lo_s = uw::smem_u32(a_lo), b_s = uw::smem_u32(b_sm);
#pragma unroll
      for (int t = 0; t < TILES; ++t) {
#pragma unroll
        for (int kb = 0; kb < US::KBLK; ++kb) {
          const uint32_t aoff = (uint32_t)t * US::A_TILE_BYTES + (uint32_t)kb * 4096u;
          const uint64_t da_hi = uw::make_desc(a_hi_s + aoff, 4096, 128);
          const uint64_t da_lo = uw::make_desc(a_lo_s + aoff, 4096, 128);
          const uint64_t db = uw::make_desc(b_s + (uint32_t)kb * (NCOL / 4 * 128), NCOL / 4 * 128, 128);
          const uint32_t d = tmem_base + (uint32_t)(t * NCOL);
          uw::umma(d, da_hi, db, IDESC_FULL, (j | kb) ? 1u : 0u);    // hi * [hi | lo] -> cols [0, 2C)
          uw::umma(d, da_lo, db, IDESC_HALF, 1u);                    // lo * hi        -> cols [0, C)
        }
      }
      uw::umma_commit(&bars[s]);
    }
  }

  // ---- epilogue: TMEM -> atomics into dW[cs][cb][tap] ----
  if (nstage > 0) {
    const int jl = nstage - 1;
    uw::mbar_wait(&bars[jl & 1], (uint32_t)((jl >> 1) & 1));
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const int wq = warp & 3, half = warp >> 2;       // TMEM lane quarter of this warp, column half
    const int row = wq * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(wq * 32) << 16);
#pragma unroll 1
    for (int t = 0; t < TILES; ++t) {
      const int tap = tap0 + t * US::TPT + row / C;
      const int cb = row % C;
      float* dst = A.dW + (long long)cb * 27 + tap;
#pragma unroll 1
      for (int c8 = half; c8 < C / 8; c8 += 2) {
        float hi[8], lo[8];
        uw::tmem_ld8(trow + (uint32_t)(t * NCOL + c8 * 8), hi);
        uw::tmem_ld8(trow + (uint32_t)(t * NCOL + C + c8 * 8), lo);
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        if (A.debug && blockIdx.x == 0 && blockIdx.y == 0 && t == 0 && c8 == 0 && row < 2)
          printf("uw dbg: row %d tap %d cb %d hi %g %g lo %g %g\n", row, tap, cb, hi[0], hi[1], lo[0], lo[1]);
        if (tap < 27) {
#pragma unroll
          for (int i = 0; i < 8; ++i) atomicAdd(dst + (long long)(c8 * 8 + i) * C * 27, hi[i] + lo[i]);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base),
                 "r"((uint32_t)US::TMEM_COLS)
                 : "memory");
  }
}

template <int C>
static int launch_umma_wgrad(UwArgs A, cudaStream_t st) {
  using US = UwShape<C>;
  auto kern = umma_wgrad_kernel<C>;
  static bool attr_done = false;
  if (!attr_done) {
    NAS3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)US::SMEM));
    attr_done = true;
  }
  // split-K: one CTA per SM (its TMEM allocation and shared memory fill the SM), chunks of at
  // least 8 stages so the TMEM -> atomics epilogue (27*C*C adds per split) stays a minor term
  long long splits = kNumSMs / US::GROUPS;
  const long long min_chunk = 8 * US::KB;
  if (splits * min_chunk > A.nvox) splits = (A.nvox + min_chunk - 1) / min_chunk;
  if (splits < 1) splits = 1;
  long long chunk = (A.nvox + splits - 1) / splits;
  chunk = (chunk + US::KB - 1) / US::KB * US::KB;
  splits = (A.nvox + chunk - 1) / chunk;
  A.chunk = (int)chunk;
  kern<<<dim3((unsigned)splits, US::GROUPS), US::THREADS, US::SMEM, st>>>(A);
  return launched("umma_wgrad");
}

// returns NAS3D_ERR_UNSUPPORTED (no error text) when the shape is not covered / the path is off
int umma_wgrad(const nas3d_conv_desc* d, const float* small, const float* big, float* dW, cudaStream_t st) {
  if (!g_opt.umma_wgrad || d->k != 3 || d->depthwise || d->Cb != d->Cs) return NAS3D_ERR_UNSUPPORTED;
  if (d->Cb != 16 && d->Cb != 32 && d->Cb != 64) return NAS3D_ERR_UNSUPPORTED;
  if (d->Cb < g_opt.umma_wgrad_min_c) return NAS3D_ERR_UNSUPPORTED;
  if (d->ld_big % 4 || d->ld_small % 4 || !aligned16(big) || !aligned16(small)) return NAS3D_ERR_UNSUPPORTED;
  UwArgs A;
  A.big = big; A.small = small; A.dW = dW;
  A.Db = d->Db; A.Hb = d->Hb; A.Wb = d->Wb; A.ldb = d->ld_big;
  A.Ds = d->Ds; A.Hs = d->Hs; A.Ws = d->Ws; A.lds = d->ld_small;
  A.stride = d->stride; A.dil = d->dil; A.pad = d->pad;
  A.nvox = (long long)d->N * d->Ds * d->Hs * d->Ws;
  A.chunk = 0;
  A.debug = g_opt.umma_wgrad_debug;
  if (A.nvox >= (1ll << 31)) return NAS3D_ERR_UNSUPPORTED;
  switch (d->Cb) {
    case 16: return launch_umma_wgrad<16>(A, st);
    case 32: return launch_umma_wgrad<32>(A, st);
    default: return launch_umma_wgrad<64>(A, st);
  }
}

}  // namespace nas3d
