// tcgen05 / TMEM weight gradient of the WIDE dense 3x3x3 convolutions (C = 16, 32, 64; any stride /
// dilation, Conv3d and ConvTranspose3d through the conv view):
//
//     dW[cs][cb][tap] = sum over small voxels o   small[o][cs] * big[o*stride - pad + tap*dil][cb]
//
// as ONE split-K GEMM per launch:   D[m = (tap, cb)][n = cs]  +=  A[m][k = o] * B[n][k = o]
//
// * M = 128 rows = 128/C taps x C big-channels per tile; a CTA owns TWO tiles (16 / 8 / 4 taps,
//   gridDim.y = 2 / 4 / 7 tap groups) and a chunk of small voxels (gridDim.x: split-K).  Its
//   accumulators stay in TMEM for the whole chunk (fp32, written ONCE); the epilogue adds them to
//   dW with atomics.
// * kind::tf32 takes K-MAJOR operands only (tools/micro/umma_layout_probe.cu,
//   profiles/r2h_*: every MN-major descriptor returns D = 0 for 32-bit operands), but NDHWC memory is
//   MN-major for this GEMM (channels contiguous, K = voxels strided).  The transposition happens in
//   registers on the way to shared memory: a thread loads 4 channels x 4 consecutive voxels (four
//   128-bit loads), and stores one 128-bit vector per CHANNEL (its 4 voxels) into the canonical
//   K-major SWIZZLE_128B layout (row = 128 B = 32 voxels, 16-byte chunk index XOR row % 8; verified
//   by the same probe, profiles/r2i_*): the 8 lanes of a store phase hold the 8 voxel quads of one
//   row, so every store phase covers all 8 bank groups - no padding, no conflicts.
// * 3xTF32 as in conv_umma.cu: x = hi + lo (hi = top 19 bits); B rows = [small_hi | small_lo]:
//       D[:, 0:2C] += A_hi * B'        D[:, 0:C] += A_lo * B_hi       epilogue: D[:, c] + D[:, C + c]
// * 32 voxels per smem stage, double buffered (registers -> smem while the previous stage's MMAs
//   run), one elected thread issues 16 MMAs per stage (K = 8 blocks advance the descriptor start
//   by 32 B inside the swizzle atom), tcgen05.commit -> mbarrier releases a stage.
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
// SmemDescriptor (sm_100), K-major SWIZZLE_128B: start>>4 [0,14) | LBO (ignored) [16,30) |
// SBO>>4 = 1024 B between 8-row groups [32,46) | version 1 [46,48) | layout 2 = SWIZZLE_128B [61,64)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// InstrDescriptor: c_format F32 (1) [4,6) | a_format TF32 (2) [7,10) | b_format TF32 (2) [10,13)
// | a, b K-major (0) [15], [16] | n>>3 [17,23) | m>>4 [24,29), M = 128
__device__ __forceinline__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
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
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
// one channel of 4 consecutive voxels -> its K-major row chunk, hi and lo parts
__device__ __forceinline__ void store_hi_lo(unsigned char* hi_base, unsigned char* lo_base, uint32_t off,
                                            float a, float b, float c, float d) {
  const float4 hi = make_float4(tf32_hi(a), tf32_hi(b), tf32_hi(c), tf32_hi(d));
  *reinterpret_cast<float4*>(hi_base + off) = hi;
  *reinterpret_cast<float4*>(lo_base + off) = make_float4(a - hi.x, b - hi.y, c - hi.z, d - hi.w);
}
// byte offset of (row r, voxel quad vq) in a K-major SWIZZLE_128B operand
__device__ __forceinline__ uint32_t sw128_off(int r, int vq) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((vq ^ (r & 7)) << 4));
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
  float* partial;          // [splits][27][C][C] (tap, cb, cs) partial sums of the K splits
  unsigned long long mw, mh, md;   // ceil(2^40 / Ws | Hs | Ds): exact n / d = (n * m) >> 40 for n * d < 2^40
};

template <int C>
struct UwShape {
  static constexpr int PRODUCERS = 512;                      // one 4-voxel x 4-channel block per producer thread
  static constexpr int THREADS = PRODUCERS + 32;             // + the MMA-issue warp
  static constexpr int KB = 32;                              // voxels per smem stage = one swizzle row
  static constexpr int CQ = C / 4;                           // 16-byte channel groups per voxel
  static constexpr int TPT = 128 / C;                        // taps per 128-row tile
  static constexpr int TILES = 2;                            // tiles per CTA
  static constexpr int TAPS_CTA = TILES * TPT;               // 16 / 8 / 4 tap slots per CTA
  static constexpr int GROUPS = (27 + TAPS_CTA - 1) / TAPS_CTA;    // gridDim.y: 2 / 4 / 7
  static constexpr int NCOL = 2 * C;                         // [hi | lo] columns of one tile
  static constexpr int TMEM_COLS = TILES * NCOL < 32 ? 32 : TILES * NCOL;    // 64 / 128 / 256
  static constexpr int A_TILE_BYTES = 128 * 128;             // 128 rows x 32 voxels, one of hi / lo
  static constexpr int A_BYTES = TILES * A_TILE_BYTES;
  static constexpr int B_BYTES = NCOL * 128;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + B_BYTES;  // multiple of 1024: every operand base is atom-aligned
  static constexpr int NST = C == 64 ? 2 : 3;                // smem stages (68 / 72 / 80 KB each)
  static constexpr size_t SMEM = (size_t)NST * STAGE_BYTES + 128 + 1024;     // + barriers + slack to align to 1024
  static_assert(TAPS_CTA * CQ * 8 == PRODUCERS, "64 (tap slot, channel quad) pairs x 8 voxel quads");
  static_assert(STAGE_BYTES % 1024 == 0, "stage must keep 1024-byte alignment");
  static_assert(SMEM <= 227 * 1024, "shared memory");
};

namespace uw {
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
}  // namespace uw

// How this kernel got its shape (measured, profiles/r2k_*, r2n_*, r2p..r2s ablations): the first
// versions ran every stage as lock-step phases between CTA barriers - all warps store, all warps
// load, one warp issues the MMAs - and the phases ADDED UP (stores 660 + loads 1100 + index math
// 1100 + MMA 300 + barrier skeleton 500 clk per 32-voxel stage; tensor pipe 4.5 % busy).  Now there
// is NO CTA barrier in the main loop:
//   warps 0..15  producers.  Per stage: registers (loaded one stage ahead) -> hi/lo split -> swizzled
//                smem, proxy fence, one mbarrier arrive per warp on full[s]; then the gather of the
//                next stage goes in flight.  Each warp decodes the 32 voxels of a stage itself (one
//                voxel per lane: multiply-shift division by host-computed reciprocals, a 10-bit "tap
//                offset stays inside the volume" mask) and hands them out with shuffles, so producers
//                never wait for each other and drift into different phases, which is what overlaps
//                global loads, shared-memory stores and the tensor core.
//   warp  16     MMA issuer: waits full[s] (16 arrivals), issues the 16 MMAs of the stage,
//                tcgen05.commit -> empty[s] tells the producers the buffer has drained.
// Everything that does not depend on the stage is computed once per thread (its tap, the tap's
// element offset, the four swizzled store offsets): a gathered 128-bit load costs a mask test and
// an add.
template <int C>
__global__ void __launch_bounds__(UwShape<C>::THREADS, 1) umma_wgrad_kernel(const UwArgs A) {
  using US = UwShape<C>;
  constexpr int KB = US::KB, CQ = US::CQ, TILES = US::TILES, NCOL = US::NCOL, NST = US::NST;
  extern __shared__ unsigned char smem_raw[];
  // SWIZZLE_128B atoms must sit on 1024-byte boundaries of the shared-memory address space
  unsigned char* smem = smem_raw + ((1024u - (uw::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NST * US::STAGE_BYTES);        // [NST]
  uint64_t* empty = full + NST;                                                        // [NST]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(empty + NST);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool producer = tid < US::PRODUCERS;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) {
      uw::mbar_init(&full[s], US::PRODUCERS / 32);
      uw::mbar_init(&empty[s], 1);
    }
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

  if (producer) {
    // ---- per-thread constants ----
    const int vq = tid & 7, pr = tid >> 3;                    // my voxel quad, my (tap slot, quad) pair
    const int tap = tap0 + pr / CQ, cq = pr % CQ;
    const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
    // bits 0-2 / 3-5 / 6-8: offset k*dil along d / h / w is inside `big`; bit 9: the voxel exists
    const int tapmask = tap < 27 ? ((1 << kd) | (8 << kh) | (64 << kw) | 512) : (1 << 30);   // bit 30 is never set
    const int delta = ((kd * A.dil * A.Hb + kh * A.dil) * A.Wb + kw * A.dil) * A.ldb + cq * 4;
    const bool is_b = pr < CQ;                                // these threads also move `small`
    uint32_t a_off[4], b_off[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a_off[i] = (uint32_t)(pr >> 5) * US::A_TILE_BYTES + uw::sw128_off((pr & 31) * 4 + i, vq);
      b_off[i] = uw::sw128_off(cq * 4 + i, vq);
    }

    // register buffers of TWO stages in flight: the loads of stage i+2 are issued right after stage i
    // has been written to shared memory, so a full stage of other work lies between a load and its use
    // (with one stage in flight every warp exposed the whole load latency once per stage)
    float4 xa0[4], xb0[4], xa1[4], xb1[4];   // [voxel of my quad] = 4 channels of big at my tap / of small
    auto gather = [&](int j, float4 (&xa)[4], float4 (&xb)[4]) {
      // lane l decodes voxel l of the stage: {element offset of (n, o*stride - pad) in big, mask,
      // element offset in small}
      const long long o = k0 + (long long)j * KB + lane;
      int ibase = 0, imask = 0, ismall = 0;
      if (o < k1) {
        const unsigned t0 = (unsigned)o;
        const unsigned q0 = (unsigned)(((unsigned long long)t0 * A.mw) >> 40);
        const int ow = (int)(t0 - q0 * (unsigned)A.Ws);
        const unsigned q1 = (unsigned)(((unsigned long long)q0 * A.mh) >> 40);
        const int oh = (int)(q0 - q1 * (unsigned)A.Hs);
        const unsigned q2 = (unsigned)(((unsigned long long)q1 * A.md) >> 40);
        const int od = (int)(q1 - q2 * (unsigned)A.Ds);
        const int n = (int)q2;
        const int bd = od * A.stride - A.pad, bh = oh * A.stride - A.pad, bw = ow * A.stride - A.pad;
        imask = 512;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          imask |= ((unsigned)(bd + k * A.dil) < (unsigned)A.Db) ? (1 << k) : 0;
          imask |= ((unsigned)(bh + k * A.dil) < (unsigned)A.Hb) ? (8 << k) : 0;
          imask |= ((unsigned)(bw + k * A.dil) < (unsigned)A.Wb) ? (64 << k) : 0;
        }
        ibase = (((n * A.Db + bd) * A.Hb + bh) * A.Wb + bw) * A.ldb;
        ismall = (int)o * A.lds;
      }
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int src = vq * 4 + v;
        const int m = __shfl_sync(0xffffffffu, imask, src);
        const int bo = __shfl_sync(0xffffffffu, ibase, src);
        const int so = __shfl_sync(0xffffffffu, ismall, src);
        const bool ok = (m & tapmask) == tapmask;
        xa[v] = uw::ldg4_pred(A.big + (ok ? bo + delta : 0), ok);
        if (is_b) xb[v] = uw::ldg4_pred(A.small + so + cq * 4, (m & 512) != 0);
      }
    };
    auto put = [&](int i, const float4 (&xa)[4], const float4 (&xb)[4]) {
      const int s = i % NST, use = i / NST;
      unsigned char* a_hi = smem + s * US::STAGE_BYTES;
      unsigned char* a_lo = a_hi + US::A_BYTES;
      unsigned char* b_sm = a_hi + 2 * US::A_BYTES;
      if (use >= 1) uw::mbar_wait(&empty[s], (uint32_t)((use - 1) & 1));     // MMAs of stage i-NST drained
      // A: rows (tap slot, cb) of my tile; one 16-byte chunk (my 4 voxels) per channel
      uw::store_hi_lo(a_hi, a_lo, a_off[0], xa[0].x, xa[1].x, xa[2].x, xa[3].x);
      uw::store_hi_lo(a_hi, a_lo, a_off[1], xa[0].y, xa[1].y, xa[2].y, xa[3].y);
      uw::store_hi_lo(a_hi, a_lo, a_off[2], xa[0].z, xa[1].z, xa[2].z, xa[3].z);
      uw::store_hi_lo(a_hi, a_lo, a_off[3], xa[0].w, xa[1].w, xa[2].w, xa[3].w);
      // B: rows [0, C) = hi(small[., cs]), rows [C, 2C) = lo
      if (is_b) {
        unsigned char* b_lo = b_sm + C * 128;
        uw::store_hi_lo(b_sm, b_lo, b_off[0], xb[0].x, xb[1].x, xb[2].x, xb[3].x);
        uw::store_hi_lo(b_sm, b_lo, b_off[1], xb[0].y, xb[1].y, xb[2].y, xb[3].y);
        uw::store_hi_lo(b_sm, b_lo, b_off[2], xb[0].z, xb[1].z, xb[2].z, xb[3].z);
        uw::store_hi_lo(b_sm, b_lo, b_off[3], xb[0].w, xb[1].w, xb[2].w, xb[3].w);
      }
      // generic-proxy stores -> visible to the tensor core, then this warp's arrival on full[s].  The
      // fence comes BEFORE the next gather: it is a memory barrier for the thread and would wait for
      // those loads
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) uw::mbar_arrive(&full[s]);
    };

    if (nstage > 0) gather(0, xa0, xb0);
    if (nstage > 1) gather(1, xa1, xb1);
#pragma unroll 1
    for (int i = 0; i < nstage; i += 2) {
      put(i, xa0, xb0);
      if (i + 2 < nstage) gather(i + 2, xa0, xb0);
      if (i + 1 < nstage) {
        put(i + 1, xa1, xb1);
        if (i + 3 < nstage) gather(i + 3, xa1, xb1);
      }
    }
  } else if (lane == 0) {
    // ---- MMA issuer ----
    constexpr uint32_t IDESC_FULL = uw::make_idesc(NCOL);
    constexpr uint32_t IDESC_HALF = uw::make_idesc(C);
#pragma unroll 1
    for (int j = 0; j < nstage; ++j) {
      const int s = j % NST;
      uw::mbar_wait(&full[s], (uint32_t)((j / NST) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const uint32_t a_hi_s = uw::smem_u32(smem + s * US::STAGE_BYTES);
      const uint32_t a_lo_s = a_hi_s + US::A_BYTES, b_s = a_hi_s + 2 * US::A_BYTES;
#pragma unroll
      for (int kb = 0; kb < KB / 8; ++kb) {
        const uint64_t db = uw::make_desc_sw128(b_s + kb * 32);
#pragma unroll
        for (int t = 0; t < TILES; ++t) {
          const uint64_t da_hi = uw::make_desc_sw128(a_hi_s + t * US::A_TILE_BYTES + kb * 32);
          const uint64_t da_lo = uw::make_desc_sw128(a_lo_s + t * US::A_TILE_BYTES + kb * 32);
          const uint32_t d = tmem_base + (uint32_t)(t * NCOL);
          uw::umma(d, da_hi, db, IDESC_FULL, (j | kb) ? 1u : 0u);    // hi * [hi | lo] -> cols [0, 2C)
          uw::umma(d, da_lo, db, IDESC_HALF, 1u);                    // lo * hi        -> cols [0, C)
        }
      }
      uw::umma_commit(&empty[s]);
    }
  }

  // ---- epilogue: TMEM -> this split's slab of the partial buffer, [tap][cb][cs] with cs contiguous
  // (plain 128-bit stores; 294 CTAs adding into the same 27*C*C addresses with atomics serialise in
  // the L2).  Warp w reads TMEM lanes 32*(w%4).., the four warps sharing a lane quarter split the
  // 8-column groups.  nstage >= 1: the host sizes the grid so that every split owns voxels. ----
  if (producer) {
    const int jl = nstage - 1;
    uw::mbar_wait(&empty[jl % NST], (uint32_t)((jl / NST) & 1));
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const int wq = warp & 3, part = warp >> 2;       // part 0..3
    const int row = wq * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(wq * 32) << 16);
    float* slab = A.partial + (long long)blockIdx.x * 27 * C * C;
#pragma unroll 1
    for (int u = part; u < TILES * (C / 8); u += 4) {
      const int t = u / (C / 8), c8 = u % (C / 8);
      const int etap = tap0 + t * US::TPT + row / C;
      const int cb = row % C;
      float hi[8], lo[8];
      uw::tmem_ld8(trow + (uint32_t)(t * NCOL + c8 * 8), hi);
      uw::tmem_ld8(trow + (uint32_t)(t * NCOL + C + c8 * 8), lo);
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      if (etap < 27) {
        float* dst = slab + ((long long)etap * C + cb) * C + c8 * 8;
        st4(dst, make_float4(hi[0] + lo[0], hi[1] + lo[1], hi[2] + lo[2], hi[3] + lo[3]));
        st4(dst + 4, make_float4(hi[4] + lo[4], hi[5] + lo[5], hi[6] + lo[6], hi[7] + lo[7]));
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

// second pass: dW[cs][cb][tap] += sum over the K splits, in split order (deterministic; dW has this
// kernel as its only writer).  One thread per (tap, cb, cs), cs fastest: coalesced slab reads.
__global__ void __launch_bounds__(256)
    umma_wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int C, float* __restrict__ dW) {
  const int total = 27 * C * C;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  // fixed summation order: 4 interleaved partial sums (independent loads in flight), combined at the end
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  const float* p = partial + i;
  int s = 0;
#pragma unroll 2
  for (; s + 4 <= splits; s += 4) {
    a0 += __ldg(p + (long long)(s + 0) * total);
    a1 += __ldg(p + (long long)(s + 1) * total);
    a2 += __ldg(p + (long long)(s + 2) * total);
    a3 += __ldg(p + (long long)(s + 3) * total);
  }
  for (; s < splits; ++s) a0 += __ldg(p + (long long)s * total);
  const float acc = (a0 + a1) + (a2 + a3);
  const int cs = i % C, cb = (i / C) % C, tap = i / (C * C);
  dW[((long long)cs * C + cb) * 27 + tap] += acc;
}

template <int C>
static long long umma_wgrad_splits(long long nvox, int* chunk_out) {
  using US = UwShape<C>;
  // split-K: about two CTAs per SM over all tap groups (one is resident per SM: shared memory),
  // chunks of at least 4 stages
  long long splits = (2 * kNumSMs + US::GROUPS - 1) / US::GROUPS;
  const long long min_chunk = 4 * US::KB;
  if (splits * min_chunk > nvox) splits = (nvox + min_chunk - 1) / min_chunk;
  if (splits < 1) splits = 1;
  long long chunk = (nvox + splits - 1) / splits;
  chunk = (chunk + US::KB - 1) / US::KB * US::KB;
  splits = (nvox + chunk - 1) / chunk;
  if (chunk_out) *chunk_out = (int)chunk;
  return splits;
}

template <int C>
static int launch_umma_wgrad(UwArgs A, long long ws_floats, cudaStream_t st) {
  using US = UwShape<C>;
  auto kern = umma_wgrad_kernel<C>;
  static bool attr_done = false;
  if (!attr_done) {
    NAS3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)US::SMEM));
    attr_done = true;
  }
  const long long splits = umma_wgrad_splits<C>(A.nvox, &A.chunk);
  NAS3D_REQUIRE(ws_floats >= splits * 27 * C * C, "conv_wgrad: workspace of %lld floats, need %lld",
                ws_floats, splits * 27 * C * C);
  kern<<<dim3((unsigned)splits, US::GROUPS), US::THREADS, US::SMEM, st>>>(A);
  int rc = launched("umma_wgrad");
  if (rc) return rc;
  umma_wgrad_reduce_kernel<<<(27 * C * C + 255) / 256, 256, 0, st>>>(A.partial, (int)splits, C, A.dW);
  return launched("umma_wgrad_reduce");
}

static bool umma_wgrad_shape(const nas3d_conv_desc* d) {
  if (!g_opt.umma_wgrad || d->k != 3 || d->depthwise || d->Cb != d->Cs) return false;
  if (d->Cb != 16 && d->Cb != 32 && d->Cb != 64) return false;
  const long long nvox_k = (long long)d->N * d->Ds * d->Hs * d->Ws;
  // measured on B200 (tools/wgrad_micro.py, profiles/r2u_wgrad_micro.txt), tcgen05 vs the CUDA-core
  // kernels: C = 64 @8^3 x8: 38 vs 100 us; C = 32 @16^3 x8: 78 vs 102 us; C = 16 stride 2 @32^3 -> 16^3
  // x8: 48 vs 62 us; but C = 16 @32^3 x8 (262 144 voxels): 225 vs 160 us - at 16 channels the tile is
  // 8 taps wide, the im2col expansion through shared memory costs more than the FFMA kernel's halo
  // reuse.  So: from umma_wgrad_min_c channels up always, 16 channels only for small K.
  if (d->Cb < g_opt.umma_wgrad_min_c && !(d->Cb == 16 && nvox_k <= 65536)) return false;
  if (d->ld_big % 4 || d->ld_small % 4) return false;
  // 32-bit element offsets inside the kernel (incl. the -pad halo start of a voxel)
  const long long big_elems = (long long)d->N * d->Db * d->Hb * d->Wb * d->ld_big;
  const long long nvox = (long long)d->N * d->Ds * d->Hs * d->Ws;
  // ... and exact multiply-shift division of a voxel index by an extent (n * d < 2^40)
  if (d->Ws > 1024 || d->Hs > 1024 || d->Ds > 1024) return false;
  return big_elems < (1ll << 30) && nvox * d->ld_small < (1ll << 30);
}

// floats of caller-provided scratch the tcgen05 path needs for this conv (0: path not taken)
long long umma_wgrad_workspace_floats(const nas3d_conv_desc* d) {
  if (!umma_wgrad_shape(d)) return 0;
  const long long nvox = (long long)d->N * d->Ds * d->Hs * d->Ws;
  const long long per = 27ll * d->Cb * d->Cb;
  switch (d->Cb) {
    case 16: return umma_wgrad_splits<16>(nvox, nullptr) * per;
    case 32: return umma_wgrad_splits<32>(nvox, nullptr) * per;
    default: return umma_wgrad_splits<64>(nvox, nullptr) * per;
  }
}

// returns NAS3D_ERR_UNSUPPORTED (no error text) when the shape is not covered, the path is off or no
// workspace was handed in
int umma_wgrad(const nas3d_conv_desc* d, const float* small, const float* big, float* dW,
               float* workspace, long long ws_floats, cudaStream_t st) {
  if (!workspace || !umma_wgrad_shape(d) || !aligned16(big) || !aligned16(small) || !aligned16(workspace))
    return NAS3D_ERR_UNSUPPORTED;
  UwArgs A;
  A.big = big; A.small = small; A.dW = dW; A.partial = workspace;
  A.Db = d->Db; A.Hb = d->Hb; A.Wb = d->Wb; A.ldb = d->ld_big;
  A.Ds = d->Ds; A.Hs = d->Hs; A.Ws = d->Ws; A.lds = d->ld_small;
  A.stride = d->stride; A.dil = d->dil; A.pad = d->pad;
  A.nvox = (long long)d->N * d->Ds * d->Hs * d->Ws;
  A.chunk = 0;
  A.mw = ((1ull << 40) + d->Ws - 1) / d->Ws;
  A.mh = ((1ull << 40) + d->Hs - 1) / d->Hs;
  A.md = ((1ull << 40) + d->Ds - 1) / d->Ds;
  switch (d->Cb) {
    case 16: return launch_umma_wgrad<16>(A, ws_floats, st);
    case 32: return launch_umma_wgrad<32>(A, ws_floats, st);
    default: return launch_umma_wgrad<64>(A, ws_floats, st);
  }
}

}  // namespace nas3d
