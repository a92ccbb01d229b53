// tcgen05 / TMEM implicit-GEMM 3x3x3 convolution for the WIDE layers (C = 16, 32, 64), any
// stride / dilation, plain and transposed, forward and dgrad:
//
//     D[128 voxels][N = Cprod]  =  sum over k = (tap, cred)   A[voxel][k] * W[k][n]
//
// * A is never materialised in HBM: the CTA's 128 threads gather their voxel's row for the
//   taps of one K-stage straight from the NDHWC activation (zero rows where the tap falls into
//   the padding or, for transposed convs, off the stride lattice) and write it to shared memory
//   in the canonical K-major SWIZZLE_NONE core-matrix layout ([m/8][k/4][m%8][16 B]).
// * fp32-grade accuracy on the TF32 pipe (1xTF32 misses the 1e-3 gradient
//   tolerance): 3xTF32 error compensation.  x = hi + lo with hi = top 19 bits; the weights are
//   pre-split and packed as B' = [W_hi | W_lo] along N, so per 8-wide K step the issuing thread
//   launches   D[:, 0:2N] += A_hi * B'      (one read of A_hi feeds hi*hi and hi*lo)
//              D[:, 0:N ] += A_lo * W_hi
//   and the epilogue adds the two N-column halves.  fp32 accumulation in TMEM.
// * one elected thread issues tcgen05.mma (cta_group::1, kind::tf32, M=128), completion is
//   tracked with tcgen05.commit -> mbarrier, two smem stages so the gather of stage i+1 overlaps
//   the MMAs of stage i; the epilogue reads TMEM with tcgen05.ld (32 lanes x 8 columns).
#include <string.h>
#include "common.cuh"

namespace nas3d {

struct UmmaArgs {
  const float* src;
  const float* wp;     // packed weights (umma_pack_kernel)
  const float* bias;
  float* dst;
  int N;
  int Dd, Hd, Wd, ldd;     // produced tensor
  int Dr, Hr, Wr, ldr;     // reduced (gathered) tensor
  int k, stride, dil, pad;
  int bfs;                 // 0: src = big (pos = o*s - p + t*d) ; 1: src = small (pos = (o + p - t*d)/s)
  int accumulate;
  long long nvox;
  double* moments;         // optional fused GN statistics (host guarantees V % 128 == 0)
  int splits;              // split-K over the tap stages (gridDim.y); > 1: atomic epilogue
  int cls;                 // 1: stride-2 transposed gather decomposed into the 8 output parity
                           //    classes (gridDim.z = 8, nvox = voxels per class), see below
};

// Parity classes of a stride-2 / pad-1 / dil-1 transposed gather (ConvTranspose3d forward, Conv3d
// dgrad): output voxel o = 2*q + p (per axis) receives input voxel (o + 1 - k) / 2 only when
// o + 1 - k is even, i.e. p = 0: tap k = 1 (input q); p = 1: taps k = 0 (input q + 1), 2 (input q).
// Walking all 27 taps for every voxel leaves 7/8 of the gathered rows (and MMAs) zero; per class
// only (1+pd)(1+ph)(1+pw) taps exist - 27 over the 8 classes instead of 27 per voxel.
// Class c = pd*4 + ph*2 + pw.  Stage offsets of the classes in the packed weights:
__host__ __device__ __forceinline__ int cls_taps(int c) {
  return (1 + ((c >> 2) & 1)) * (1 + ((c >> 1) & 1)) * (1 + (c & 1));
}
template <int TPS>
__host__ __device__ __forceinline__ int cls_stages(int c) { return (cls_taps(c) + TPS - 1) / TPS; }
template <int TPS>
__host__ __device__ __forceinline__ int cls_stage_base(int c) {
  int b = 0;
  for (int i = 0; i < c; ++i) b += cls_stages<TPS>(i);
  return b;
}
// slot-th tap of class c -> (kd, kh, kw) and the input offsets (+1 where k = 0)
__host__ __device__ __forceinline__ void cls_tap(int c, int slot, int* kd, int* kh, int* kw,
                                                 int* od, int* oh, int* ow) {
  const int pd = (c >> 2) & 1, ph = (c >> 1) & 1, pw = c & 1;
  const int aw = slot % (1 + pw); slot /= (1 + pw);
  const int ah = slot % (1 + ph); slot /= (1 + ph);
  const int ad = slot;
  *kd = pd ? (ad ? 2 : 0) : 1; *od = (pd && !ad) ? 1 : 0;
  *kh = ph ? (ah ? 2 : 0) : 1; *oh = (ph && !ah) ? 1 : 0;
  *kw = pw ? (aw ? 2 : 0) : 1; *ow = (pw && !aw) ? 1 : 0;
}

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
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  // SmemDescriptor (sm_100): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48)
  // | layout_type = SWIZZLE_NONE (0) [61,64)
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ constexpr uint32_t make_idesc_tf32(int n) {
  // InstrDescriptor: c_format F32 (1) [4,6) | a_format TF32 (2) [7,10) | b_format TF32 (2) [10,13)
  // | a,b K-major (0) | n>>3 [17,23) | m>>4 [24,29)  with M = 128
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
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

// branch-free predicated 128-bit load (zeros when !ok) so that all gathers of a stage are in
// flight together instead of being serialised behind per-load branches
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
__device__ __forceinline__ void cp_async16_u(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc));
}

template <int CRED, int NPROD>
struct UmmaShape {
  static constexpr int KS = CRED == 16 ? 32 : CRED;      // K floats per smem stage
  static constexpr int TPS = KS / CRED;                  // taps per stage
  static constexpr int NIT = (27 + TPS - 1) / TPS;       // stages per tile
  static constexpr int KCH = KS / 4;                     // 16-byte chunks per row per stage
  static constexpr int A_BYTES = 128 * KS * 4;
  static constexpr int B_BYTES = 2 * NPROD * KS * 4;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * NPROD < 32 ? 32 : 2 * NPROD;   // 32 / 64 / 128
  static constexpr size_t SMEM = 2 * STAGE_BYTES + 64;
  static constexpr int PACKED_FLOATS = NIT * 2 * NPROD * KS;
};

// Wp[it][n/8][c][n%8][e] ; k = it*KS + c*4 + e -> (tap = k / CRED, cr = k % CRED)
// rows n < NPROD: hi(W[prod = n][cr][tap]) ; rows n >= NPROD: lo(...)
// bfs: 0 = reduce Cb (small-from-big), 1 = reduce Cs (big-from-small), 2 = big-from-small in
// parity-class stage order (stride 2, see cls_tap)
template <int CRED, int NPROD>
__device__ __forceinline__ void umma_pack_body(const float* __restrict__ w, int Cb, int bfs,
                                               float* __restrict__ wp) {
  using US = UmmaShape<CRED, NPROD>;
  const int total = US::PACKED_FLOATS;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int r = i;
    const int e = r % 4; r /= 4;
    const int n8 = r % 8; r /= 8;
    const int c = r % US::KCH; r /= US::KCH;
    const int nb = r % (2 * NPROD / 8);
    const int it = r / (2 * NPROD / 8);
    const int n = nb * 8 + n8;
    const int kk = it * US::KS + c * 4 + e;
    int tap = kk / CRED;
    const int cr = kk % CRED;
    if (bfs == 2) {
      // global stage `it` -> (class, local stage); slot = local stage * TPS + tap-in-stage
      int cl = 0, base = 0;
      while (cl < 7 && it >= base + cls_stages<US::TPS>(cl)) { base += cls_stages<US::TPS>(cl); ++cl; }
      const int slot = (it - base) * US::TPS + (c * 4 + e) / CRED;
      tap = 27;
      if (it < base + cls_stages<US::TPS>(cl) && slot < cls_taps(cl)) {
        int kd, kh, kw, o0, o1, o2;
        cls_tap(cl, slot, &kd, &kh, &kw, &o0, &o1, &o2);
        tap = (kd * 3 + kh) * 3 + kw;
      }
    }
    float v = 0.f;
    if (tap < 27) {
      const int prod = n % NPROD;
      // W[cs][cb][tap]; SFB: prod = cs, red = cb ; BFS: red = cs, prod = cb
      const long long gi = bfs ? ((long long)cr * Cb + prod) * 27 + tap
                               : ((long long)prod * Cb + cr) * 27 + tap;
      const float x = __ldg(w + gi);
      const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
      v = (n < NPROD) ? hi : (x - hi);
    }
    wp[i] = v;
  }
}

template <int CRED, int NPROD>
__global__ void umma_pack_kernel(const float* __restrict__ w, int Cb, int bfs,
                                 float* __restrict__ wp) {
  umma_pack_body<CRED, NPROD>(w, Cb, bfs, wp);
}

// all tcgen05 weight operands of a forward/backward in ONE launch (blockIdx.y = entry): they only
// change in optimizer.step, and 38 separate 8-us pack launches per step were 0.3 ms of the graph
constexpr int PACK_BATCH = 48;
struct PackBatch {
  const float* w[PACK_BATCH];
  float* wp[PACK_BATCH];
  unsigned char C[PACK_BATCH];
  unsigned char bfs[PACK_BATCH];
};
__global__ void __launch_bounds__(256) umma_pack_batch_kernel(const __grid_constant__ PackBatch B) {
  const int e = blockIdx.y;
  const int Cc = B.C[e], bfs = B.bfs[e];
  if (Cc == 16) umma_pack_body<16, 16>(B.w[e], 16, bfs, B.wp[e]);
  else if (Cc == 32) umma_pack_body<32, 32>(B.w[e], 32, bfs, B.wp[e]);
  else umma_pack_body<64, 64>(B.w[e], 64, bfs, B.wp[e]);
}

template <int CRED, int NPROD>
__global__ void __launch_bounds__(128) umma_conv_kernel(const UmmaArgs A) {
  using US = UmmaShape<CRED, NPROD>;
  constexpr int KS = US::KS, KCH = US::KCH, NIT = US::NIT;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * US::STAGE_BYTES);   // empty[0], empty[1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 2 * US::STAGE_BYTES + 32);
  __shared__ double sm_mom[2 * NPROD];
  if (threadIdx.x < 2 * NPROD) sm_mom[threadIdx.x] = 0.0;

  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"((uint32_t)US::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // my output voxel (row m = tid of the 128-row tile)
  long long o = (long long)blockIdx.x * 128 + tid;
  const bool row_valid = o < A.nvox;
  const int cl = A.cls ? (int)blockIdx.z : 0;
  int n = 0, od = 0, oh = 0, ow = 0;
  if (row_valid) {
    long long t = o;
    if (A.cls) {
      // row index inside the parity class -> output voxel 2*q + p
      const int Wq = A.Wd >> 1, Hq = A.Hd >> 1, Dq = A.Dd >> 1;
      ow = 2 * (int)(t % Wq) + (cl & 1); t /= Wq;
      oh = 2 * (int)(t % Hq) + ((cl >> 1) & 1); t /= Hq;
      od = 2 * (int)(t % Dq) + ((cl >> 2) & 1);
      n = (int)(t / Dq);
      o = (((long long)n * A.Dd + od) * A.Hd + oh) * A.Wd + ow;
    } else {
      ow = (int)(t % A.Wd); t /= A.Wd;
      oh = (int)(t % A.Hd); t /= A.Hd;
      od = (int)(t % A.Dd);
      n = (int)(t / A.Dd);
    }
  }
  const float* src_n = A.src + (long long)n * A.Dr * A.Hr * A.Wr * A.ldr;
  const uint32_t row_off = (uint32_t)((tid >> 3) * (KCH * 128) + (tid & 7) * 16);   // bytes

  constexpr uint32_t IDESC1 = make_idesc_tf32(2 * NPROD);
  constexpr uint32_t IDESC2 = make_idesc_tf32(NPROD);

  // gather of one stage into registers: my row, TPS taps x CRED channels
  float4 xv[US::TPS][CRED / 4];
  auto gather = [&](int it) {
#pragma unroll
    for (int tp = 0; tp < US::TPS; ++tp) {
      const int tap = it * US::TPS + tp;
      int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
      int id, ih, iw;
      bool ok = row_valid && tap < 27;
      if (A.cls) {
        // `it` is the class-local stage, tap its slot: every row of the tile has this tap
        int o0, o1, o2;
        ok = row_valid && tap < cls_taps(cl);
        cls_tap(cl, ok ? tap : 0, &kd, &kh, &kw, &o0, &o1, &o2);
        id = (od >> 1) + o0; ih = (oh >> 1) + o1; iw = (ow >> 1) + o2;
      } else if (A.bfs) {
        const int nd = od + A.pad - kd * A.dil, nh = oh + A.pad - kh * A.dil,
                  nw = ow + A.pad - kw * A.dil;
        ok = ok && nd >= 0 && nh >= 0 && nw >= 0 && (nd % A.stride) == 0 &&
             (nh % A.stride) == 0 && (nw % A.stride) == 0;
        id = nd / A.stride; ih = nh / A.stride; iw = nw / A.stride;
      } else {
        id = od * A.stride - A.pad + kd * A.dil;
        ih = oh * A.stride - A.pad + kh * A.dil;
        iw = ow * A.stride - A.pad + kw * A.dil;
      }
      ok = ok && id >= 0 && id < A.Dr && ih >= 0 && ih < A.Hr && iw >= 0 && iw < A.Wr;
      const float* px = ok ? src_n + (((long long)id * A.Hr + ih) * A.Wr + iw) * A.ldr : A.src;
#pragma unroll
      for (int c = 0; c < CRED / 4; ++c) xv[tp][c] = ldg4_pred(px + c * 4, ok);
    }
  };
  // split-K: this CTA reduces the K stages [it0, it1) (small problems have too few voxel tiles to
  // fill 148 SMs and would serialise 27 gather->MMA round trips per tile)
  const int nit = A.cls ? cls_stages<US::TPS>(cl) : NIT;
  const int wbase = A.cls ? cls_stage_base<US::TPS>(cl) : 0;     // first packed stage of my class
  const int it0 = (int)((long long)blockIdx.y * nit / A.splits);
  const int it1 = (int)((long long)(blockIdx.y + 1) * nit / A.splits);
  gather(it0);
  // B: packed weights of a stage (contiguous), asynchronous copy - issued ONE STAGE AHEAD so that its
  // L2 latency overlaps the A stores / gather of the stage before (it used to be issued and
  // awaited inside the same iteration: ~1 us of every 1.3 us iteration, 27 iterations per tile)
  auto load_b = [&](int it, int s) {
    unsigned char* b_sm = smem + s * US::STAGE_BYTES + 2 * US::A_BYTES;
    const float* g = A.wp + (long long)(wbase + it) * (US::B_BYTES / 4);
#pragma unroll
    for (int i = tid; i < US::B_BYTES / 16; i += 128) cp_async16_u(b_sm + i * 16, g + i * 4);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  if (it0 < it1) load_b(it0, 0);

#pragma unroll 1
  for (int it = it0; it < it1; ++it) {
    const int j = it - it0;
    const int s = j & 1;
    unsigned char* stage = smem + s * US::STAGE_BYTES;
    unsigned char* a_hi = stage;
    unsigned char* a_lo = stage + US::A_BYTES;
    unsigned char* b_sm = stage + 2 * US::A_BYTES;
    // (the MMAs of stage j-2 were awaited in iteration j-1, before B(j) was sent into this stage)
    // ---- A: split the gathered row into hi / lo and store in core-matrix layout ----
#pragma unroll
    for (int tp = 0; tp < US::TPS; ++tp) {
#pragma unroll
      for (int c = 0; c < CRED / 4; ++c) {
        const float4 x = xv[tp][c];
        float4 hi, lo;
        hi.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
        hi.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
        hi.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
        hi.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
        lo.x = x.x - hi.x; lo.y = x.y - hi.y; lo.z = x.z - hi.z; lo.w = x.w - hi.w;
        const uint32_t off = row_off + (uint32_t)(tp * (CRED / 4) + c) * 128u;
        *reinterpret_cast<float4*>(a_hi + off) = hi;
        *reinterpret_cast<float4*>(a_lo + off) = lo;
      }
    }
    // next stage's gather goes in flight now; it lands while this stage's MMAs are issued
    if (it + 1 < it1) {
      gather(it + 1);
      if (j >= 1) mbar_wait(&bars[s ^ 1], (uint32_t)(((j - 1) >> 1) & 1));   // MMAs of stage j-1 drained
      load_b(it + 1, s ^ 1);
      asm volatile("cp.async.wait_group 1;\n" ::: "memory");     // B(j) has landed, B(j+1) in flight
    } else {
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    }
    // generic-proxy smem writes -> visible to the tensor core (async proxy)
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const uint32_t a_hi_s = smem_u32(a_hi), a_lo_s = smem_u32(a_lo), b_s = smem_u32(b_sm);
#pragma unroll
      for (int kk = 0; kk < KS / 8; ++kk) {
        const uint64_t da_hi = make_smem_desc(a_hi_s + kk * 256, 128, KCH * 128);
        const uint64_t da_lo = make_smem_desc(a_lo_s + kk * 256, 128, KCH * 128);
        const uint64_t db = make_smem_desc(b_s + kk * 256, 128, KCH * 128);
        umma_tf32(tmem_base, da_hi, db, IDESC1, (j | kk) ? 1u : 0u);    // hi*[hi|lo] -> cols [0,2N)
        umma_tf32(tmem_base, da_lo, db, IDESC2, 1u);                    // lo*hi      -> cols [0,N)
      }
      umma_commit(&bars[s]);
    }
  }

  // ---- epilogue: all MMAs done when the last stage's commit lands ----
  const int jl = it1 - it0 - 1;
  mbar_wait(&bars[jl & 1], (uint32_t)((jl >> 1) & 1));
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
  float* pd = A.dst + o * A.ldd;
#pragma unroll 1
  for (int c8 = 0; c8 < NPROD / 8; ++c8) {
    float hi[8], lo[8];
    tmem_ld8(trow + c8 * 8, hi);
    tmem_ld8(trow + NPROD + c8 * 8, lo);
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    if (row_valid) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[i] = hi[i] + lo[i];
        if (A.bias && blockIdx.y == 0) v[i] += __ldg(A.bias + c8 * 8 + i);
      }
      if (A.splits > 1) {
        // partial sums of this K split: dst was zero-initialised (or holds the running gradient)
#pragma unroll
        for (int i = 0; i < 8; ++i) atomicAdd(pd + c8 * 8 + i, v[i]);
      } else {
        float4 v0 = make_float4(v[0], v[1], v[2], v[3]), v1 = make_float4(v[4], v[5], v[6], v[7]);
        if (A.accumulate) {
          const float4 o0 = *reinterpret_cast<const float4*>(pd + c8 * 8);
          const float4 o1 = *reinterpret_cast<const float4*>(pd + c8 * 8 + 4);
          v0.x += o0.x; v0.y += o0.y; v0.z += o0.z; v0.w += o0.w;
          v1.x += o1.x; v1.y += o1.y; v1.z += o1.z; v1.w += o1.w;
        }
        st4(pd + c8 * 8, v0);
        st4(pd + c8 * 8 + 4, v1);
        if (A.moments) {
          const float s0[4] = {v0.x, v0.y, v0.z, v0.w}, s1[4] = {v1.x, v1.y, v1.z, v1.w};
          const float q0[4] = {v0.x * v0.x, v0.y * v0.y, v0.z * v0.z, v0.w * v0.w};
          const float q1[4] = {v1.x * v1.x, v1.y * v1.y, v1.z * v1.z, v1.w * v1.w};
          warp_moments_add(sm_mom, c8 * 8, s0, q0);
          warp_moments_add(sm_mom, c8 * 8 + 4, s1, q1);
        }
      }
    }
  }
  if (A.moments) {
    __syncthreads();
    cta_moments_flush(sm_mom, A.moments, n, NPROD, 128);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base),
                 "r"((uint32_t)US::TMEM_COLS)
                 : "memory");
  }
}

// =========================================================================================
// Warp-specialised variant (the default, option umma_ws): same operands, same MMA sequence, but
// the three jobs no longer take turns.
//   * warps 0..7 (256 threads) PRODUCE: two threads per voxel row, each gathers half of a
//     stage's K slice into registers ONE STAGE AHEAD (the loads of stage j+1 are issued before
//     the hi/lo split and shared-memory stores of stage j), writes its chunks of A_hi / A_lo
//     and arrives on full[s]; thread 0 also sends the stage's packed weights with one
//     cp.async.bulk onto the same barrier (expect_tx).
//   * lane 0 of warp 8 ISSUES: waits full[s], issues the stage's tcgen05.mma chain and commits it
//     to empty[s] (which the producers wait on before they reuse the stage).
//   * all 8 producer warps run the EPILOGUE: warp w reads TMEM lanes 32*(w%4) and the column half
//     w/4.
// In the lock-step kernel above every tap cost one exposed L2 round trip plus a CTA barrier
// (1.3 - 2.4 us per stage with the tensor pipe idle most of it).
// =========================================================================================
template <int CRED, int NPROD>
struct UmmaWsShape : UmmaShape<CRED, NPROD> {
  using US = UmmaShape<CRED, NPROD>;
  static constexpr int NSTG = 2;
  static constexpr int PROD = 256, THREADS = PROD + 32;
  static constexpr int HCH = US::KCH / 2;                // 16-byte chunks per producer thread and stage
  static constexpr size_t SMEM = NSTG * US::STAGE_BYTES + 128;
};

template <int CRED, int NPROD>
__global__ void __launch_bounds__(UmmaWsShape<CRED, NPROD>::THREADS, CRED == 64 ? 1 : 2)   // 2 CTAs: <= 112 registers
    umma_conv_ws_kernel(const UmmaArgs A) {
  using US = UmmaShape<CRED, NPROD>;
  using WS = UmmaWsShape<CRED, NPROD>;
  constexpr int KS = US::KS, KCH = US::KCH, NIT = US::NIT, NSTG = WS::NSTG, HCH = WS::HCH;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSTG * US::STAGE_BYTES);    // [NSTG]
  uint64_t* empty = full + NSTG;                                                   // [NSTG]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(empty + NSTG);
  __shared__ double sm_mom[2 * NPROD];
  if (threadIdx.x < 2 * NPROD) sm_mom[threadIdx.x] = 0.0;

  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    for (int b = 0; b < NSTG; ++b) {
      mbar_init(&full[b], WS::PROD + 1);     // every producer + thread 0's expect_tx arrival
      mbar_init(&empty[b], 1);               // one tcgen05.commit
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"((uint32_t)US::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  const int cl = A.cls ? (int)blockIdx.z : 0;
  const int nit = A.cls ? cls_stages<US::TPS>(cl) : NIT;
  const int wbase = A.cls ? cls_stage_base<US::TPS>(cl) : 0;     // first packed stage of my class
  const int it0 = (int)((long long)blockIdx.y * nit / A.splits);
  const int it1 = (int)((long long)(blockIdx.y + 1) * nit / A.splits);
  constexpr uint32_t IDESC1 = make_idesc_tf32(2 * NPROD);
  constexpr uint32_t IDESC2 = make_idesc_tf32(NPROD);

  if (warp == WS::PROD / 32) {
    // ---- MMA issuer ---------------------------------------------------------------------------
    if ((tid & 31) == 0) {
#pragma unroll 1
      for (int it = it0; it < it1; ++it) {
        const int j = it - it0, s = j % NSTG;
        mbar_wait(&full[s], (uint32_t)((j / NSTG) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        unsigned char* stage = smem + s * US::STAGE_BYTES;
        const uint32_t a_hi_s = smem_u32(stage), a_lo_s = smem_u32(stage + US::A_BYTES),
                       b_s = smem_u32(stage + 2 * US::A_BYTES);
#pragma unroll
        for (int kk = 0; kk < KS / 8; ++kk) {
          const uint64_t da_hi = make_smem_desc(a_hi_s + kk * 256, 128, KCH * 128);
          const uint64_t da_lo = make_smem_desc(a_lo_s + kk * 256, 128, KCH * 128);
          const uint64_t db = make_smem_desc(b_s + kk * 256, 128, KCH * 128);
          umma_tf32(tmem_base, da_hi, db, IDESC1, (j | kk) ? 1u : 0u);    // hi*[hi|lo] -> cols [0,2N)
          umma_tf32(tmem_base, da_lo, db, IDESC2, 1u);                    // lo*hi      -> cols [0,N)
        }
        umma_commit(&empty[s]);
      }
    }
  } else {
    // ---- producers ----------------------------------------------------------------------------
    const int row = tid & 127, half = tid >> 7;
    long long o = (long long)blockIdx.x * 128 + row;
    const bool row_valid = o < A.nvox;
    int n = 0, od = 0, oh = 0, ow = 0;
    if (row_valid) {
      long long t = o;
      if (A.cls) {
        const int Wq = A.Wd >> 1, Hq = A.Hd >> 1, Dq = A.Dd >> 1;
        ow = 2 * (int)(t % Wq) + (cl & 1); t /= Wq;
        oh = 2 * (int)(t % Hq) + ((cl >> 1) & 1); t /= Hq;
        od = 2 * (int)(t % Dq) + ((cl >> 2) & 1);
        n = (int)(t / Dq);
        o = (((long long)n * A.Dd + od) * A.Hd + oh) * A.Wd + ow;
      } else {
        ow = (int)(t % A.Wd); t /= A.Wd;
        oh = (int)(t % A.Hd); t /= A.Hd;
        od = (int)(t % A.Dd);
        n = (int)(t / A.Dd);
      }
    }
    const float* src_n = A.src + (long long)n * A.Dr * A.Hr * A.Wr * A.ldr;
    const uint32_t row_off = (uint32_t)((row >> 3) * (KCH * 128) + (row & 7) * 16);   // bytes

    // my half of a stage's K slice: chunks [half*HCH, (half+1)*HCH) = floats k0 .. k0 + 4*HCH of the
    // stage; one tap (TPS = 1: half the channels of it; TPS = 2: tap `half`, all channels).
    // Everything that does not depend on the tap is computed once per thread: the input coordinate of
    // tap k along an axis is (base + sgn * k * dil) >> sh, valid when non-negative, on the stride
    // lattice and inside the tensor; offsets are 32-bit (host-checked).  ncu r4o: ~200 instructions
    // per thread and stage, most of them this arithmetic in 64 bits with runtime divisions.
    const int sgn = A.bfs ? -1 : 1, sh = A.bfs ? A.stride - 1 : 0, step = sgn * A.dil;
    const int bd = A.bfs ? od + A.pad : od * A.stride - A.pad;
    const int bh = A.bfs ? oh + A.pad : oh * A.stride - A.pad;
    const int bw = A.bfs ? ow + A.pad : ow * A.stride - A.pad;
    const int qd = od >> 1, qh = oh >> 1, qw = ow >> 1;        // parity-class mode
    const int c0 = US::TPS == 2 ? 0 : half * (CRED / 2);        // first channel of my chunks
    auto gather = [&](int it, float4 (&xv)[HCH]) {
      const int tap = it * US::TPS + (US::TPS == 2 ? half : 0);
      int id, ih, iw;
      bool ok;
      if (A.cls) {
        int kd, kh, kw, o0, o1, o2;
        ok = row_valid && tap < cls_taps(cl);
        cls_tap(cl, ok ? tap : 0, &kd, &kh, &kw, &o0, &o1, &o2);
        id = qd + o0; ih = qh + o1; iw = qw + o2;
      } else {
        const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
        const int nd = bd + kd * step, nh = bh + kh * step, nw = bw + kw * step;
        ok = row_valid && tap < 27 && (nd | nh | nw) >= 0 && ((nd | nh | nw) & sh) == 0;
        id = nd >> sh; ih = nh >> sh; iw = nw >> sh;
      }
      ok = ok && id >= 0 && id < A.Dr && ih >= 0 && ih < A.Hr && iw >= 0 && iw < A.Wr;
      const unsigned off = (unsigned)(((id * A.Hr + ih) * A.Wr + iw) * A.ldr + c0);
      const float* px = ok ? src_n + off : A.src;
#pragma unroll
      for (int c = 0; c < HCH; ++c) xv[c] = ldg4_pred(px + c * 4, ok);
    };
    // stage j: wait for its smem, (thread 0) send B, split + store my chunks, arrive
    auto produce = [&](int it, const float4 (&xv)[HCH]) {
      const int j = it - it0, s = j % NSTG;
      unsigned char* stage = smem + s * US::STAGE_BYTES;
      if (j >= NSTG) mbar_wait(&empty[s], (uint32_t)(((j / NSTG) - 1) & 1));   // MMAs of stage j-NSTG drained
      if (tid == 0) {
        const uint32_t bar = smem_u32(&full[s]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar),
                     "r"((uint32_t)US::B_BYTES)
                     : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                smem_u32(stage + 2 * US::A_BYTES)),
            "l"(A.wp + (long long)(wbase + it) * (US::B_BYTES / 4)), "r"((uint32_t)US::B_BYTES), "r"(bar)
            : "memory");
      }
#pragma unroll
      for (int c = 0; c < HCH; ++c) {
        const float4 x = xv[c];
        float4 hi, lo;
        hi.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
        hi.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
        hi.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
        hi.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
        lo.x = x.x - hi.x; lo.y = x.y - hi.y; lo.z = x.z - hi.z; lo.w = x.w - hi.w;
        const uint32_t off = row_off + (uint32_t)(half * HCH + c) * 128u;
        *reinterpret_cast<float4*>(stage + off) = hi;
        *reinterpret_cast<float4*>(stage + US::A_BYTES + off) = lo;
      }
      // generic-proxy smem writes -> visible to the tensor core (async proxy), then hand over
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(&full[s])) : "memory");
    };

    // three register sets in rotation: the loads of stages j+1 and j+2 are in flight while stage j is
    // split and stored (ncu r4o: 32-47 % of the producers' time was the gather's long scoreboard with
    // one stage of look-ahead)
    float4 xa[HCH], xb[HCH], xc[HCH];
    if (it0 < it1) gather(it0, xa);
    if (it0 + 1 < it1) gather(it0 + 1, xb);
#pragma unroll 1
    for (int it = it0; it < it1; it += 3) {
      if (it + 2 < it1) gather(it + 2, xc);
      produce(it, xa);
      if (it + 1 < it1) {
        if (it + 3 < it1) gather(it + 3, xa);
        produce(it + 1, xb);
      }
      if (it + 2 < it1) {
        if (it + 4 < it1) gather(it + 4, xb);
        produce(it + 2, xc);
      }
    }

    // ---- epilogue: the last stage's commit covers every MMA before it -------------------------
    const int jl = it1 - it0 - 1;
    mbar_wait(&empty[jl % NSTG], (uint32_t)((jl / NSTG) & 1));
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    // warp w: TMEM lanes 32*(w%4) (row = 32*(w%4) + lane), column half w/4
    const int erow = (warp & 3) * 32 + (tid & 31), ehalf = warp >> 2;
    long long eo = (long long)blockIdx.x * 128 + erow;
    const bool erow_valid = eo < A.nvox;
    int en = 0;
    if (erow_valid) {
      long long t = eo;
      if (A.cls) {
        const int Wq = A.Wd >> 1, Hq = A.Hd >> 1, Dq = A.Dd >> 1;
        const int w_ = 2 * (int)(t % Wq) + (cl & 1); t /= Wq;
        const int h_ = 2 * (int)(t % Hq) + ((cl >> 1) & 1); t /= Hq;
        const int d_ = 2 * (int)(t % Dq) + ((cl >> 2) & 1);
        en = (int)(t / Dq);
        eo = (((long long)en * A.Dd + d_) * A.Hd + h_) * A.Wd + w_;
      } else {
        en = (int)(t / ((long long)A.Dd * A.Hd * A.Wd));
      }
    }
    n = en;
    const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    float* pd = A.dst + eo * A.ldd;
    constexpr int C8H = NPROD / 16;          // 8-column groups per half
#pragma unroll 1
    for (int c8 = ehalf * C8H; c8 < (ehalf + 1) * C8H; ++c8) {
      float hi[8], lo[8];
      tmem_ld8(trow + c8 * 8, hi);
      tmem_ld8(trow + NPROD + c8 * 8, lo);
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      if (erow_valid) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          v[i] = hi[i] + lo[i];
          if (A.bias && blockIdx.y == 0) v[i] += __ldg(A.bias + c8 * 8 + i);
        }
        if (A.splits > 1) {
#pragma unroll
          for (int i = 0; i < 8; ++i) atomicAdd(pd + c8 * 8 + i, v[i]);
        } else {
          float4 v0 = make_float4(v[0], v[1], v[2], v[3]), v1 = make_float4(v[4], v[5], v[6], v[7]);
          if (A.accumulate) {
            const float4 o0 = *reinterpret_cast<const float4*>(pd + c8 * 8);
            const float4 o1 = *reinterpret_cast<const float4*>(pd + c8 * 8 + 4);
            v0.x += o0.x; v0.y += o0.y; v0.z += o0.z; v0.w += o0.w;
            v1.x += o1.x; v1.y += o1.y; v1.z += o1.z; v1.w += o1.w;
          }
          st4(pd + c8 * 8, v0);
          st4(pd + c8 * 8 + 4, v1);
          if (A.moments) {
            const float s0[4] = {v0.x, v0.y, v0.z, v0.w}, s1[4] = {v1.x, v1.y, v1.z, v1.w};
            const float q0[4] = {v0.x * v0.x, v0.y * v0.y, v0.z * v0.z, v0.w * v0.w};
            const float q1[4] = {v1.x * v1.x, v1.y * v1.y, v1.z * v1.z, v1.w * v1.w};
            warp_moments_add(sm_mom, c8 * 8, s0, q0);
            warp_moments_add(sm_mom, c8 * 8 + 4, s1, q1);
          }
        }
      } else if (A.moments && A.splits == 1) {
        // keep the warp-collective reductions convergent (host guarantees full tiles with moments)
        const float z[4] = {0.f, 0.f, 0.f, 0.f};
        warp_moments_add(sm_mom, c8 * 8, z, z);
        warp_moments_add(sm_mom, c8 * 8 + 4, z, z);
      }
    }
    if (A.moments) {
      // producers only (named barrier 1): the issuer warp has nothing to add
      asm volatile("bar.sync 1, %0;\n" ::"n"(WS::PROD) : "memory");
      for (int i = tid; i < 2 * NPROD; i += WS::PROD)
        atomicAdd(&A.moments[(long long)n * NPROD * 2 + i], sm_mom[i]);
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

template <int CRED, int NPROD>
static int launch_umma(const UmmaArgs& A, cudaStream_t st) {
  using US = UmmaShape<CRED, NPROD>;
  auto kern = umma_conv_kernel<CRED, NPROD>;
  static bool attr_done = false;
  if (!attr_done) {
    NAS3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)US::SMEM));
    attr_done = true;
  }
  const unsigned blocks = (unsigned)((A.nvox + 127) / 128);
  // (its gather shifts instead of dividing and keeps in-sample offsets in 32 bits)
  // Measured (profiles/r4r_umma_direct.txt): faster or equal everywhere except the 16-channel parity-class
  // dgrad (42 vs 33 us at 8 x 32^3), which keeps the lock-step kernel.
  if (g_opt.umma_ws && (A.stride == 1 || A.stride == 2) && !(A.cls && CRED == 16) &&
      (long long)A.Dr * A.Hr * A.Wr * A.ldr < (1ll << 31)) {
    using WS = UmmaWsShape<CRED, NPROD>;
    auto kws = umma_conv_ws_kernel<CRED, NPROD>;
    static bool ws_attr_done = false;
    if (!ws_attr_done) {
      NAS3D_CUDA(cudaFuncSetAttribute(kws, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS::SMEM));
      ws_attr_done = true;
    }
    kws<<<dim3(blocks, (unsigned)A.splits, A.cls ? 8u : 1u), WS::THREADS, WS::SMEM, st>>>(A);
    return launched("umma_conv_ws");
  }
  kern<<<dim3(blocks, (unsigned)A.splits, A.cls ? 8u : 1u), 128, US::SMEM, st>>>(A);
  return launched("umma_conv");
}

template <int CRED, int NPROD>
static int launch_pack(const float* w, int Cb, int bfs, float* wp, cudaStream_t st) {
  using US = UmmaShape<CRED, NPROD>;
  umma_pack_kernel<CRED, NPROD><<<(US::PACKED_FLOATS + 255) / 256, 256, 0, st>>>(w, Cb, bfs, wp);
  return launched("umma_pack");
}

static bool umma_channels_ok(int c) { return c == 16 || c == 32 || c == 64; }

// stride-2 transposed gather with even big extents: parity-class decomposition
static bool umma_class_mode(const nas3d_conv_desc* d, int produce_big) {
  return produce_big && d->k == 3 && d->stride == 2 && d->dil == 1 && d->pad == 1 &&
         d->Db == 2 * d->Ds && d->Hb == 2 * d->Hs && d->Wb == 2 * d->Ws;
}

}  // namespace nas3d

using namespace nas3d;

extern "C" {

long long nas3d_umma_packed_floats(const nas3d_conv_desc* d, int produce_big) {
  if (!d || d->k != 3 || d->depthwise || d->Cb != d->Cs || !umma_channels_ok(d->Cb)) return 0;
  (void)produce_big;
  switch (d->Cb) {
    case 16: return UmmaShape<16, 16>::PACKED_FLOATS;
    case 32: return UmmaShape<32, 32>::PACKED_FLOATS;
    default: return UmmaShape<64, 64>::PACKED_FLOATS;
  }
}

int nas3d_umma_pack_mode(const nas3d_conv_desc* d, int produce_big) {
  if (!d || !produce_big) return 0;
  return umma_class_mode(d, produce_big) ? 2 : 1;
}

int nas3d_umma_pack_weights(const nas3d_conv_desc* d, const float* w, int produce_big, float* packed,
                            void* stream) {
  NAS3D_REQUIRE(nas3d_umma_packed_floats(d, produce_big) > 0, "umma_pack: unsupported conv shape");
  NAS3D_REQUIRE(aligned16(packed), "umma_pack: packed buffer must be 16B aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int mode = nas3d_umma_pack_mode(d, produce_big);
  switch (d->Cb) {
    case 16: return launch_pack<16, 16>(w, d->Cb, mode, packed, st);
    case 32: return launch_pack<32, 32>(w, d->Cb, mode, packed, st);
    default: return launch_pack<64, 64>(w, d->Cb, mode, packed, st);
  }
}

int nas3d_umma_pack_weights_batch(int n, const int* channels, const int* modes,
                                  const float* const* w, float* const* packed, void* stream) {
  NAS3D_REQUIRE(n >= 0, "umma_pack_batch: n=%d", n);
  cudaStream_t st = (cudaStream_t)stream;
  for (int base = 0; base < n; base += PACK_BATCH) {
    const int m = n - base < PACK_BATCH ? n - base : PACK_BATCH;
    PackBatch B;
    memset(&B, 0, sizeof(B));
    int cmax = 16;
    for (int i = 0; i < m; ++i) {
      const int c = channels[base + i], md = modes[base + i];
      NAS3D_REQUIRE(umma_channels_ok(c) && md >= 0 && md <= 2, "umma_pack_batch: entry %d: C=%d mode=%d", base + i, c, md);
      NAS3D_REQUIRE(aligned16(packed[base + i]), "umma_pack_batch: packed buffer must be 16B aligned");
      B.w[i] = w[base + i]; B.wp[i] = packed[base + i];
      B.C[i] = (unsigned char)c; B.bfs[i] = (unsigned char)md;
      if (c > cmax) cmax = c;
    }
    const int gx = cmax == 16 ? 28 : cmax == 32 ? 108 : 216;   // ~2-4 grid-stride passes of the widest entry
    umma_pack_batch_kernel<<<dim3(gx, m), 256, 0, st>>>(B);
    int rc = launched("umma_pack_batch");
    if (rc) return rc;
  }
  return NAS3D_OK;
}

int nas3d_umma_conv(const nas3d_conv_desc* d, int produce_big, const float* src,
                    const float* packed_w, const float* bias, float* dst, int accumulate,
                    double* moments, void* stream) {
  NAS3D_REQUIRE(nas3d_umma_packed_floats(d, produce_big) > 0, "umma_conv: unsupported conv shape");
  NAS3D_REQUIRE(d->ld_big % 4 == 0 && d->ld_small % 4 == 0 && aligned16(src) && aligned16(dst) &&
                    aligned16(packed_w),
                "umma_conv: pitches must be multiples of 4 and pointers 16B aligned");
  for (int ax = 0; ax < 3; ++ax) {
    const int b = ax == 0 ? d->Db : ax == 1 ? d->Hb : d->Wb;
    const int s = ax == 0 ? d->Ds : ax == 1 ? d->Hs : d->Ws;
    const int expect = (b + 2 * d->pad - d->dil * (d->k - 1) - 1) / d->stride + 1;
    NAS3D_REQUIRE(s == expect, "umma_conv: small extent %d != %d implied by big extent %d", s, expect, b);
  }
  UmmaArgs A;
  A.src = src; A.wp = packed_w; A.bias = bias; A.dst = dst; A.N = d->N;
  if (produce_big) {
    A.Dd = d->Db; A.Hd = d->Hb; A.Wd = d->Wb; A.ldd = d->ld_big;
    A.Dr = d->Ds; A.Hr = d->Hs; A.Wr = d->Ws; A.ldr = d->ld_small;
  } else {
    A.Dd = d->Ds; A.Hd = d->Hs; A.Wd = d->Ws; A.ldd = d->ld_small;
    A.Dr = d->Db; A.Hr = d->Hb; A.Wr = d->Wb; A.ldr = d->ld_big;
  }
  A.k = d->k; A.stride = d->stride; A.dil = d->dil; A.pad = d->pad;
  A.bfs = produce_big ? 1 : 0;
  A.accumulate = accumulate;
  A.nvox = (long long)d->N * A.Dd * A.Hd * A.Wd;
  A.cls = umma_class_mode(d, produce_big) ? 1 : 0;
  if (A.cls) A.nvox /= 8;                       // rows per parity class (gridDim.z = 8)
  const long long V = (long long)A.Dd * A.Hd * A.Wd;
  cudaStream_t st = (cudaStream_t)stream;
  // split-K when the voxel tiles alone cannot fill the machine (dense destination required)
  const long long tiles = (A.nvox + 127) / 128;
  const int Cprod = produce_big ? d->Cb : d->Cs;
  int splits = 1;
  if (g_opt.umma_split_k && !A.cls && tiles < kNumSMs && A.ldd == Cprod) {
    splits = (int)(kNumSMs / tiles);
    if (splits > 9) splits = 9;
    if (splits < 1) splits = 1;
  }
  A.splits = splits;
  if (splits > 1 && !accumulate)
    NAS3D_CUDA(cudaMemsetAsync(dst, 0, sizeof(float) * (size_t)A.nvox * Cprod, st));
  const bool fuse = moments != nullptr && (A.cls ? (V / 8) % 128 == 0 : V % 128 == 0) && splits == 1;
  A.moments = fuse ? moments : nullptr;
  if (fuse)
    NAS3D_CUDA(cudaMemsetAsync(moments, 0, sizeof(double) * 2 * (size_t)d->N * (produce_big ? d->Cb : d->Cs), st));
  int rc;
  switch (d->Cb) {
    case 16: rc = launch_umma<16, 16>(A, st); break;
    case 32: rc = launch_umma<32, 32>(A, st); break;
    default: rc = launch_umma<64, 64>(A, st); break;
  }
  if (rc == NAS3D_OK && moments && !fuse)     // tile straddles samples: separate statistics pass
    rc = nas3d_moments_nc(dst, d->N, V, produce_big ? d->Cb : d->Cs, A.ldd, moments, stream);
  return rc;
}

}  // extern "C"
