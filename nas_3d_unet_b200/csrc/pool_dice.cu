// 2x2x2 pooling (prim_ops.py:160-168) and the multi-region soft-Dice loss (loss.py:12-14).
// Both are HBM-bound streaming kernels.
#include "common.cuh"

namespace nas3d {

// one thread per (output voxel, 4-channel group)
template <bool MAX>
__global__ void __launch_bounds__(256)
    pool2_fwd_kernel(const float* __restrict__ x, int ld_x, float* __restrict__ y, int ld_y,
                     int Do, int Ho, int Wo, int C4, long long total) {
  const int Hi = Ho * 2, Wi = Wo * 2, Di = Do * 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long o = i / C4;
    const int c = (int)(i - o * C4) * 4;
    long long t = o;
    const int ow = (int)(t % Wo); t /= Wo;
    const int oh = (int)(t % Ho); t /= Ho;
    const int od = (int)(t % Do);
    const long long n = t / Do;
    float4 acc = MAX ? make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY)
                     : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int dd = 0; dd < 2; ++dd)
#pragma unroll
      for (int dh = 0; dh < 2; ++dh)
#pragma unroll
        for (int dw = 0; dw < 2; ++dw) {
          const float4 v = ldg4(
              x + (((n * Di + od * 2 + dd) * Hi + oh * 2 + dh) * Wi + ow * 2 + dw) * ld_x + c);
          if (MAX) {
            // torch: (val > max) || isnan(val)
            acc.x = (v.x > acc.x || v.x != v.x) ? v.x : acc.x;
            acc.y = (v.y > acc.y || v.y != v.y) ? v.y : acc.y;
            acc.z = (v.z > acc.z || v.z != v.z) ? v.z : acc.z;
            acc.w = (v.w > acc.w || v.w != v.w) ? v.w : acc.w;
          } else {
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
          }
        }
    if (!MAX) { acc.x *= 0.125f; acc.y *= 0.125f; acc.z *= 0.125f; acc.w *= 0.125f; }
    st4(y + o * ld_y + c, acc);
  }
}

template <bool MAX>
__global__ void __launch_bounds__(256)
    pool2_bwd_kernel(const float* __restrict__ x, int ld_x, const float* __restrict__ dy, int ld_dy,
                     float* __restrict__ dx, int ld_dx, int accumulate, int Do, int Ho, int Wo,
                     int C4, long long total) {
  const int Hi = Ho * 2, Wi = Wo * 2, Di = Do * 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long o = i / C4;
    const int c = (int)(i - o * C4) * 4;
    long long t = o;
    const int ow = (int)(t % Wo); t /= Wo;
    const int oh = (int)(t % Ho); t /= Ho;
    const int od = (int)(t % Do);
    const long long n = t / Do;
    const float4 g = ldg4(dy + o * ld_dy + c);
    int arg[4] = {0, 0, 0, 0};
    if (MAX) {
      float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int dd = j >> 2, dh = (j >> 1) & 1, dw = j & 1;
        const float4 v4 = ldg4(
            x + (((n * Di + od * 2 + dd) * Hi + oh * 2 + dh) * Wi + ow * 2 + dw) * ld_x + c);
        const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (v[e] > m[e] || v[e] != v[e]) { m[e] = v[e]; arg[e] = j; }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int dd = j >> 2, dh = (j >> 1) & 1, dw = j & 1;
      float4 r;
      if (MAX) {
        r.x = arg[0] == j ? g.x : 0.f; r.y = arg[1] == j ? g.y : 0.f;
        r.z = arg[2] == j ? g.z : 0.f; r.w = arg[3] == j ? g.w : 0.f;
      } else {
        r = make_float4(g.x * 0.125f, g.y * 0.125f, g.z * 0.125f, g.w * 0.125f);
      }
      float* pd =
          dx + (((n * Di + od * 2 + dd) * Hi + oh * 2 + dh) * Wi + ow * 2 + dw) * ld_dx + c;
      if (accumulate) {
        const float4 o4 = *reinterpret_cast<const float4*>(pd);
        r.x += o4.x; r.y += o4.y; r.z += o4.z; r.w += o4.w;
      }
      st4(pd, r);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Dice: sums[n][c] = {sum p*t, sum p, sum t}; a thread walks voxels, all (<=4) channels of a
// channel group at once so NDHWC predictions are read with full sector use.
// ---------------------------------------------------------------------------------------
constexpr int DB = 256;
constexpr int DITER = 16;

__device__ __forceinline__ float ld_truth(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_truth(const signed char* p) { return (float)__ldg(p); }

template <typename TT>
__global__ void __launch_bounds__(DB)
    dice_sums_kernel(const float* __restrict__ pred, long long p_sn, long long p_sc, long long p_sv,
                     const TT* __restrict__ truth, long long t_sn, long long t_sc,
                     long long t_sv, int C, long long V, double* __restrict__ sums) {
  const int n = blockIdx.y;
  const int c0 = blockIdx.z * 4;
  const int cn = min(4, C - c0);
  float s[4][3];
#pragma unroll
  for (int e = 0; e < 4; ++e) s[e][0] = s[e][1] = s[e][2] = 0.f;
  const float* pb = pred + n * p_sn + c0 * p_sc;
  const TT* tb = truth + n * t_sn + c0 * t_sc;
  for (int it = 0; it < DITER; ++it) {
    const long long v = ((long long)blockIdx.x * DITER + it) * DB + threadIdx.x;
    if (v >= V) break;
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (e < cn) {
        const float p = __ldg(pb + v * p_sv + e * p_sc);
        const float t = ld_truth(tb + v * t_sv + e * t_sc);
        s[e][0] += p * t; s[e][1] += p; s[e][2] += t;
      }
  }
  __shared__ double red[DB / 32][12];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double v = warp_sum((double)s[e][j]);
      if (lane == 0) red[wid][e * 3 + j] = v;
    }
  __syncthreads();
  if (threadIdx.x < cn * 3) {
    double v = 0.0;
    for (int w = 0; w < DB / 32; ++w) v += red[w][threadIdx.x];
    atomicAdd(&sums[((long long)n * C + c0) * 3 + threadIdx.x], v);
  }
}

__global__ void dice_finish_kernel(const double* __restrict__ sums, int NC, float smooth,
                                   float* __restrict__ loss) {
  // single warp; NC is tiny (N*3)
  double part = 0.0;
  for (int i = threadIdx.x; i < NC; i += 32) {
    // same association order as loss.py:13-14, in fp32 like the reference
    const float I = (float)sums[i * 3], P = (float)sums[i * 3 + 1], T = (float)sums[i * 3 + 2];
    part += (double)((2.f * I + smooth) / (P + T + smooth));
  }
  part = warp_sum(part);
  if (threadIdx.x == 0) loss[0] = 1.f - (float)(part / NC);
}

template <typename TT>
__global__ void __launch_bounds__(DB)
    dice_bwd_kernel(const double* __restrict__ sums, const float* __restrict__ gout,
                    const TT* __restrict__ truth, long long t_sn, long long t_sc, long long t_sv,
                    float* __restrict__ dpred, long long p_sn, long long p_sc, long long p_sv,
                    int N, int C, long long V, float smooth) {
  const int n = blockIdx.y;
  const int c0 = blockIdx.z * 4;
  const int cn = min(4, C - c0);
  __shared__ float cA[4], cB[4];
  if (threadIdx.x < cn) {
    const double* s = sums + ((long long)n * C + c0 + threadIdx.x) * 3;
    const double D = s[1] + s[2] + (double)smooth;
    const double g = (double)gout[0] / ((double)N * C);
    cA[threadIdx.x] = (float)(-g * 2.0 / D);
    cB[threadIdx.x] = (float)(g * (2.0 * s[0] + (double)smooth) / (D * D));
  }
  __syncthreads();
  const TT* tb = truth + n * t_sn + c0 * t_sc;
  float* pb = dpred + n * p_sn + c0 * p_sc;
  for (int it = 0; it < DITER; ++it) {
    const long long v = ((long long)blockIdx.x * DITER + it) * DB + threadIdx.x;
    if (v >= V) break;
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (e < cn) pb[v * p_sv + e * p_sc] = cA[e] * ld_truth(tb + v * t_sv + e * t_sc) + cB[e];
  }
}

static inline unsigned cap_grid(long long total) {
  long long b = (total + 255) / 256;
  long long cap = (long long)kNumSMs * 8;
  return (unsigned)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace nas3d

using namespace nas3d;

extern "C" {

int nas3d_pool2_fwd(int kind, const float* x, int ld_x, float* y, int ld_y, int N, int Do, int Ho,
                    int Wo, int C, void* stream) {
  NAS3D_REQUIRE(C % 4 == 0 && ld_x % 4 == 0 && ld_y % 4 == 0 && aligned16(x) && aligned16(y),
                "pool2_fwd: C=%d and pitches must be multiples of 4", C);
  NAS3D_REQUIRE(kind == 0 || kind == 1, "pool2_fwd: kind %d", kind);
  const int C4 = C / 4;
  const long long total = (long long)N * Do * Ho * Wo * C4;
  if (kind == 1)
    pool2_fwd_kernel<true><<<cap_grid(total), 256, 0, (cudaStream_t)stream>>>(x, ld_x, y, ld_y, Do,
                                                                              Ho, Wo, C4, total);
  else
    pool2_fwd_kernel<false><<<cap_grid(total), 256, 0, (cudaStream_t)stream>>>(x, ld_x, y, ld_y, Do,
                                                                               Ho, Wo, C4, total);
  return launched("pool2_fwd");
}

int nas3d_pool2_bwd(int kind, const float* x, int ld_x, const float* dy, int ld_dy, float* dx,
                    int ld_dx, int accumulate, int N, int Do, int Ho, int Wo, int C, void* stream) {
  NAS3D_REQUIRE(C % 4 == 0 && ld_dy % 4 == 0 && ld_dx % 4 == 0 && aligned16(dy) && aligned16(dx),
                "pool2_bwd: C=%d and pitches must be multiples of 4", C);
  NAS3D_REQUIRE(kind == 0 || (kind == 1 && x != nullptr && ld_x % 4 == 0), "pool2_bwd: kind %d", kind);
  const int C4 = C / 4;
  const long long total = (long long)N * Do * Ho * Wo * C4;
  if (kind == 1)
    pool2_bwd_kernel<true><<<cap_grid(total), 256, 0, (cudaStream_t)stream>>>(
        x, ld_x, dy, ld_dy, dx, ld_dx, accumulate, Do, Ho, Wo, C4, total);
  else
    pool2_bwd_kernel<false><<<cap_grid(total), 256, 0, (cudaStream_t)stream>>>(
        x, ld_x, dy, ld_dy, dx, ld_dx, accumulate, Do, Ho, Wo, C4, total);
  return launched("pool2_bwd");
}

int nas3d_dice_fwd(const float* pred, long long p_sn, long long p_sc, long long p_sv,
                   const void* truth, int truth_is_int8, long long t_sn, long long t_sc,
                   long long t_sv, int N, int C, long long V, float smooth, double* sums, float* loss,
                   void* stream) {
  NAS3D_REQUIRE(N > 0 && C > 0 && V > 0, "dice_fwd: empty input");
  cudaStream_t st = (cudaStream_t)stream;
  NAS3D_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 3 * (size_t)N * C, st));
  dim3 grid((unsigned)((V + DB * DITER - 1) / (DB * DITER)), N, (C + 3) / 4);
  if (truth_is_int8)
    dice_sums_kernel<signed char><<<grid, DB, 0, st>>>(pred, p_sn, p_sc, p_sv, (const signed char*)truth,
                                                       t_sn, t_sc, t_sv, C, V, sums);
  else
    dice_sums_kernel<float><<<grid, DB, 0, st>>>(pred, p_sn, p_sc, p_sv, (const float*)truth, t_sn,
                                                 t_sc, t_sv, C, V, sums);
  int rc = launched("dice_sums");
  if (rc) return rc;
  dice_finish_kernel<<<1, 32, 0, st>>>(sums, N * C, smooth, loss);
  return launched("dice_finish");
}

int nas3d_dice_bwd(const double* sums, const float* gout, const void* truth, int truth_is_int8,
                   long long t_sn, long long t_sc, long long t_sv, float* dpred, long long p_sn,
                   long long p_sc, long long p_sv, int N, int C, long long V, float smooth,
                   void* stream) {
  dim3 grid((unsigned)((V + DB * DITER - 1) / (DB * DITER)), N, (C + 3) / 4);
  if (truth_is_int8)
    dice_bwd_kernel<signed char><<<grid, DB, 0, (cudaStream_t)stream>>>(
        sums, gout, (const signed char*)truth, t_sn, t_sc, t_sv, dpred, p_sn, p_sc, p_sv, N, C, V, smooth);
  else
    dice_bwd_kernel<float><<<grid, DB, 0, (cudaStream_t)stream>>>(
        sums, gout, (const float*)truth, t_sn, t_sc, t_sv, dpred, p_sn, p_sc, p_sv, N, C, V, smooth);
  return launched("dice_bwd");
}

}  // extern "C"
