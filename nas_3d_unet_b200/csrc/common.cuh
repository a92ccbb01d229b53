// Shared helpers for the nas3d_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include "../../include/nas3d_b200.h"

namespace nas3d {

extern thread_local char g_err[512];
extern std::atomic<unsigned long long> g_launches;

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// call after every kernel launch
inline int launched(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
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

// odd part / power-of-two part of the float4-group count of a channel dimension
static inline void split_c4(int C4, int* U, int* P) {
  int u = C4, p = 1;
  while ((u & 1) == 0) { u >>= 1; p <<= 1; }
  *U = u; *P = p;
}

}  // namespace nas3d
