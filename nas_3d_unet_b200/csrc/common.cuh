// Shared helpers for the nas3d_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdarg.h>
#include <atomic>
#include "../../include/nas3d_b200.h"

namespace nas3d {

extern thread_local char g_err[512];
extern std::atomic<unsigned long long> g_launches;

// kernel-selection options (options.cu): read once at load, changed only via nas3d_set_option
struct Options {
  int tiled = 1;          // specialised tiled / pointwise conv kernels (0: generic gather kernels)
  int tma = 1;            // TMA halo staging in conv3_s1 (0: cp.async staging)
  int tma_merged = 1;     // merged (W,C) inner dimension for dense C=4 TMA tiles
  int affine_ring = 1;    // cp.async-ring variants of the streaming kernels (big, narrow tensors)
  int apply_ring = 1;
  int reduce_ring = 1;
  int pw_fwd_ring = 1;
  int reduce_waves = 1;   // whole-wave grid sizing of the backward reductions
  int ring_min_log2 = 22; // smallest tensor (log2 float4 elements) the ring kernels take
  int pw_vpt_sfb = 4, pw_vpt_bfs = 2, pw_vpt_mom = 2;   // voxels per thread of the 1x1 kernels
  int umma_split_k = 1;   // tcgen05 fwd / dgrad: split the taps over CTAs (fp32 atomics) when the voxel tiles alone
                          // cannot fill the machine; 0 = one CTA walks all taps of its tile (bit-reproducible)
  int umma_wgrad = 1;     // tcgen05 weight gradient of the wide dense 3x3x3 convs (conv_umma_wgrad.cu)
  int umma_wgrad_min_c = 32;   // ... always from this channel count up; 16 channels only for small K
  int umma_ws = 1;        // warp-specialised tcgen05 fwd / dgrad kernel (0: the lock-step kernel)
  int s1_wgrad_tma = 1;   // TMA tile ring for the stride-1 weight gradient at C = 4 / 8 (0: cp.async kernel)
  int s2_wgrad_tma = 1;   // TMA-staged double-buffered stride-2 weight gradient (0: cp.async kernel)
};
extern Options g_opt;

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// per-kernel-variant launch counters (nas3d_launch_count_of): keyed by the address of the label
// literal, so the launch path is a short pointer scan; labels with equal text are summed on read
struct VariantCounter {
  std::atomic<const char*> name;
  std::atomic<unsigned long long> n;
};
constexpr int kMaxVariants = 192;
extern VariantCounter g_variants[kMaxVariants];

inline void count_variant(const char* what) {
  for (int i = 0; i < kMaxVariants; ++i) {
    const char* cur = g_variants[i].name.load(std::memory_order_relaxed);
    if (cur == nullptr) {
      const char* expect = nullptr;
      if (g_variants[i].name.compare_exchange_strong(expect, what)) cur = what;
      else cur = expect;
    }
    if (cur == what) {
      g_variants[i].n.fetch_add(1, std::memory_order_relaxed);
      return;
    }
  }
}

// call after every kernel launch
inline int launched(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  count_variant(what);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(NAS3D_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return NAS3D_OK;
}

#define NAS3D_CUDA(call)                                                              \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess)                                                           \
      return nas3d::fail(NAS3D_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__));   \
  } while (0)

#define NAS3D_REQUIRE(cond, ...)                                  \
  do {                                                            \
    if (!(cond)) return nas3d::fail(NAS3D_ERR_ARG, __VA_ARGS__);  \
  } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

constexpr int kNumSMs = 148;  // B200

__device__ __forceinline__ float4 ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// Packed fp32 FMA (sm_100 FFMA2, PTX fma.rn.f32x2): TWO fused multiply-adds per issued
// instruction.  Measured on B200 (tools/micro/fma_patterns.cu, profiles/r1e_fma_patterns_micro.jsonl):
// with the operand patterns of these kernels both forms sustain the same 58-61 TFLOP/s
// (~100 of 128 FMA/clk/SM), so the gain is in ISSUE SLOTS, not pipe throughput: the loads and
// address arithmetic of the conv loops now share the scheduler with half as many FMA
// instructions (+10 % on the fwd / dgrad kernels).  ptxas folds the {s,s} operand into the
// instruction's scalar-broadcast modifier (FFMA2 Rd, Rs.F32, Rw.F32x2.HI_LO, Rd.F32x2.HI_LO),
// no extra MOV.  Each half is an IEEE fma.rn, so results are bit-identical to two fmaf() in the
// same order.  Build with -DNAS3D_NO_FFMA2 for the scalar form (A/B measurements).
__device__ __forceinline__ void fma2(float2& d, float s, float wx, float wy) {
#ifndef NAS3D_NO_FFMA2
  asm("{\n\t"
      ".reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %2};\n\t"
      "mov.b64 rb, {%3, %4};\n\t"
      "mov.b64 rd, {%0, %1};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rd;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t"
      "}\n"
      : "+f"(d.x), "+f"(d.y)
      : "f"(s), "f"(wx), "f"(wy));
#else
  d.x = fmaf(s, wx, d.x);
  d.y = fmaf(s, wy, d.y);
#endif
}
// element-wise pair: d += a * b
__device__ __forceinline__ void fma2v(float2& d, float ax, float ay, float bx, float by) {
#ifndef NAS3D_NO_FFMA2
  asm("{\n\t"
      ".reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %5};\n\t"
      "mov.b64 rd, {%0, %1};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rd;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t"
      "}\n"
      : "+f"(d.x), "+f"(d.y)
      : "f"(ax), "f"(ay), "f"(bx), "f"(by));
#else
  d.x = fmaf(ax, bx, d.x);
  d.y = fmaf(ay, by, d.y);
#endif
}
// 4 accumulators (two pairs) += s * w
__device__ __forceinline__ void axpy4(float2 (&a)[2], float s, const float4& w) {
  fma2(a[0], s, w.x, w.y);
  fma2(a[1], s, w.z, w.w);
}
// 4x4 outer product: a[i][0..3] += x[i] * w[0..3]   (wgrad: x = activation, w = dy)
__device__ __forceinline__ void outer4(float2 (&a)[4][2], const float4& x, const float4& w) {
  axpy4(a[0], x.x, w);
  axpy4(a[1], x.y, w);
  axpy4(a[2], x.z, w);
  axpy4(a[3], x.w, w);
}
// 4-channel contraction: a[0..3] += sum_i x[i] * w[i][0..3]   (fwd / dgrad)
__device__ __forceinline__ void dot4x4(float2 (&a)[2], const float4& x, const float4 (&w)[4]) {
  axpy4(a, x.x, w[0]);
  axpy4(a, x.y, w[1]);
  axpy4(a, x.z, w[2]);
  axpy4(a, x.w, w[3]);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Warp "reduce-scatter" of N per-lane values (N a multiple of 32): on return lane l holds in
// v[0 .. N/32) the warp totals of the original v[l*(N/32) + j].  Each of the 5 butterfly rounds
// exchanges only the half a lane does not keep: N*(31/32) shuffles in total instead of 5*N for N
// separate all-reduces.
template <int N>
__device__ __forceinline__ void warp_reduce_scatter(float (&v)[N]) {
  static_assert(N % 32 == 0, "warp_reduce_scatter: N must be a multiple of 32");
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < 5; ++r) {
    const int o = 16 >> r;
    const int n = N >> (r + 1);
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const float send = up ? v[i] : v[i + n];
      const float keep = up ? v[i + n] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
}

// Fused GroupNorm statistics for conv epilogues: a warp whose 32 lanes hold values of the SAME 4
// channels (c0..c0+3) of one sample reduces its per-lane partial {sum, sum of squares} and adds
// them to a per-CTA shared array sm[2*C] (doubles); cta_moments_flush() then does one fp64
// atomicAdd per (channel, moment) per CTA into moments[n][C][2].
__device__ __forceinline__ void warp_moments_add(double* sm, int c0, const float s[4],
                                                 const float q[4]) {
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const double a = warp_sum((double)s[e]);
    const double b = warp_sum((double)q[e]);
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&sm[(c0 + e) * 2 + 0], a);
      atomicAdd(&sm[(c0 + e) * 2 + 1], b);
    }
  }
}
__device__ __forceinline__ void cta_moments_flush(const double* sm, double* moments, int n, int C,
                                                  int nthreads) {
  for (int i = threadIdx.x; i < 2 * C; i += nthreads)
    atomicAdd(&moments[(long long)n * C * 2 + i], sm[i]);
}

// smallest tensor (float4 elements per launch) the ring-staged streaming kernels take: 2^22 is what
// was measured (batch 8 at 128^3 / 64^3); option ring_min_log2
static inline long long ring_min_elems() { return 1ll << g_opt.ring_min_log2; }

// odd part / power-of-two part of the float4-group count of a channel dimension
static inline void split_c4(int C4, int* U, int* P) {
  int u = C4, p = 1;
  while ((u & 1) == 0) { u >>= 1; p <<= 1; }
  *U = u; *P = p;
}

}  // namespace nas3d
