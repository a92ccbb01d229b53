// Sliding-window inference chain on the device: what prediction.py:121-170 and
// patches.py:99-206 do on the host with numpy, bit-exact for every integer / float64 result.
//   extract_patches : volume (C,D,H,W) + patch corners  -> NDHWC patch batch, zero outside
//                     (get_patch_from_3d_data + fix_out_of_bound_patch_attempt, patches.py:99-169)
//   stitch_labels   : patch predictions -> float64 mean over overlaps accumulated IN PATCH ORDER
//                     (stitch, patches.py:172-206), threshold + label assembly (get_tumor_pred,
//                     prediction.py:150-170), skull mask, paste at the brain offset
//   seg_to_masks    : int16 segmentation -> 3 fp32 region masks (get_multi_class_labels,
//                     generator.py:230-248, including its WT quirk)
// HBM-bound gathers; one thread per output voxel.
#include "common.cuh"

namespace nas3d {

__global__ void __launch_bounds__(256)
    extract_patches_kernel(const float* __restrict__ vol, int C, int D, int H, int W,
                           const int* __restrict__ corners, int B, int Pd, int Ph, int Pw,
                           float* __restrict__ out, int ld, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int pw = (int)(t % Pw); t /= Pw;
    const int ph = (int)(t % Ph); t /= Ph;
    const int pd = (int)(t % Pd);
    const int b = (int)(t / Pd);
    const int gd = corners[b * 3 + 0] + pd, gh = corners[b * 3 + 1] + ph, gw = corners[b * 3 + 2] + pw;
    const bool ok = gd >= 0 && gd < D && gh >= 0 && gh < H && gw >= 0 && gw < W;
    float* o = out + i * ld;
    const long long v = ((long long)gd * H + gh) * W + gw;
    for (int c = 0; c < C; ++c) o[c] = ok ? __ldg(vol + (long long)c * D * H * W + v) : 0.f;
    for (int c = C; c < ld; ++c) o[c] = 0.f;
  }
}

// patch_nonzero[b] = 1 iff patch b of the volume holds any non-zero voxel (prediction.py:133: an
// all-zero patch is not run through the net, its prediction is DEFINED as zeros).  One CTA-strided
// pass; flags must be zeroed by the caller.
__global__ void __launch_bounds__(256)
    patch_nonzero_kernel(const float* __restrict__ vol, int C, int D, int H, int W,
                         const int* __restrict__ corners, int Pd, int Ph, int Pw,
                         int* __restrict__ flags, long long per_patch) {
  const int b = blockIdx.y;
  const int c0 = corners[b * 3 + 0], c1 = corners[b * 3 + 1], c2 = corners[b * 3 + 2];
  bool any = false;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_patch && !any;
       i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int pw = (int)(t % Pw); t /= Pw;
    const int ph = (int)(t % Ph);
    const int pd = (int)(t / Ph);
    const int gd = c0 + pd, gh = c1 + ph, gw = c2 + pw;
    if (gd < 0 || gd >= D || gh < 0 || gh >= H || gw < 0 || gw >= W) continue;
    const long long v = ((long long)gd * H + gh) * W + gw;
    for (int c = 0; c < C; ++c) any |= (__ldg(vol + (long long)c * D * H * W + v) != 0.f);
  }
  if (__syncthreads_or(any ? 1 : 0) && threadIdx.x == 0) atomicOr(flags + b, 1);
}

struct StitchArgs {
  const float* const* preds;   // [B] device table: patch b = [Pd][Ph][Pw][ldp], 3 channels used
  const int* nonzero;          // [B] or NULL: 0 = the patch prediction counts as all zeros
  const int* corners;      // [B][3] in brain coordinates (may be negative)
  int B, Pd, Ph, Pw, ldp;
  int Db, Hb, Wb;          // brain box
  int D, H, W;             // full volume
  int od, oh, ow;          // brain box offset in the full volume (brain_width[0])
  float threshold;
  int inclusive;
  const unsigned char* skull;   // [D][H][W] or NULL
  unsigned char* labels;        // [D][H][W]
  double* stitched;             // optional [3][Db][Hb][Wb]
};

__global__ void __launch_bounds__(256) stitch_labels_kernel(const StitchArgs A, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int w = (int)(t % A.W); t /= A.W;
    const int h = (int)(t % A.H);
    const int d = (int)(t / A.H);
    const int bd = d - A.od, bh = h - A.oh, bw = w - A.ow;
    unsigned char lab = 0;
    if (bd >= 0 && bd < A.Db && bh >= 0 && bh < A.Hb && bw >= 0 && bw < A.Wb) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0;
      int cnt = 0;
      for (int b = 0; b < A.B; ++b) {       // patch order == numpy's accumulation order
        const int pd = bd - A.corners[b * 3 + 0], ph = bh - A.corners[b * 3 + 1],
                  pw = bw - A.corners[b * 3 + 2];
        if (pd < 0 || pd >= A.Pd || ph < 0 || ph >= A.Ph || pw < 0 || pw >= A.Pw) continue;
        ++cnt;
        if (A.nonzero && !A.nonzero[b]) continue;       // += 0.0: an exact no-op on the sums
        const float* p = A.preds[b] + (((long long)pd * A.Ph + ph) * A.Pw + pw) * A.ldp;
        s0 += (double)__ldg(p); s1 += (double)__ldg(p + 1); s2 += (double)__ldg(p + 2);
      }
      const double c = cnt ? (double)cnt : 1.0;
      const double m0 = s0 / c, m1 = s1 / c, m2 = s2 / c;
      if (A.stitched) {
        const long long bv = ((long long)bd * A.Hb + bh) * A.Wb + bw;
        const long long plane = (long long)A.Db * A.Hb * A.Wb;
        A.stitched[bv] = m0; A.stitched[plane + bv] = m1; A.stitched[2 * plane + bv] = m2;
      }
      const double th = (double)A.threshold;
      const bool a = m0 >= th, bb = m1 >= th, cc = m2 >= th;
      if (A.inclusive) {
        if (bb) lab = 2;
        if (a) lab = 1;
        if (cc) lab = 4;
      } else {
        int tv = (a ? 1 : 0) + (bb ? 2 : 0);
        if (tv == 3) tv = (m0 >= m1) ? 1 : 2;
        if (cc) tv += 4;
        if (tv == 5) tv = (m0 >= m2) ? 1 : 4;
        if (tv == 6) tv = (m1 >= m2) ? 2 : 4;
        lab = (unsigned char)tv;
      }
    }
    if (A.skull) lab = (unsigned char)(lab * A.skull[i]);
    A.labels[i] = lab;
  }
}

__global__ void __launch_bounds__(256)
    seg_to_masks_kernel(const short* __restrict__ seg, long long V, long long total, int inclusive,
                        float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / V, v = i - n * V;
    const short t = seg[i];
    float c0, c1, c2;
    if (inclusive) {
      c0 = (t == 1 || t == 4) ? 1.f : 0.f;
      c1 = (t == 1 || t == 2) ? 1.f : 0.f;      // the reference's logical_or(.., .., out) quirk
      c2 = (t == 4) ? 1.f : 0.f;
    } else {
      c0 = (t == 1) ? 1.f : 0.f; c1 = (t == 2) ? 1.f : 0.f; c2 = (t == 4) ? 1.f : 0.f;
    }
    float* o = out + n * 3 * V + v;
    o[0] = c0; o[V] = c1; o[2 * V] = c2;
  }
}

// Batch staging after the H2D copy (generator.py:195-248): per-sample cube permutation (one of the
// 48 keys of augment.permute_data, given as a signed-stride index map) of the 4-modality patch into
// the NDHWC layout the stem reads, and int16 segmentation -> three int8 region masks.
constexpr int kStageMaxN = 64;
struct StageArgs {
  const float* x;        // [N][C][V] planar (what np.asarray(x_list) uploads)
  const void* seg;       // [N][V] int16 (as stored) or int8 (values 0,1,2,4 fit: half the PCIe bytes), or NULL
  int seg_bytes;         // 2 | 1
  float* xo;             // [N][V][ld]
  signed char* yo;       // [N][3][V] or NULL
  int N, C, D, H, W, ld, inclusive;
  int base[kStageMaxN];
  int sd[kStageMaxN][3];
};

__global__ void __launch_bounds__(256) stage_patches_kernel(const __grid_constant__ StageArgs a) {
  const long long V = (long long)a.D * a.H * a.W;
  const long long total = V * a.N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / V);
    long long v = i - (long long)n * V;
    const int w = (int)(v % a.W);
    const int h = (int)((v / a.W) % a.H);
    const int d = (int)(v / ((long long)a.W * a.H));
    const long long src = (long long)a.base[n] + (long long)a.sd[n][0] * d +
                          (long long)a.sd[n][1] * h + (long long)a.sd[n][2] * w;
    float* o = a.xo + i * a.ld;
    const float* xs = a.x + (long long)n * a.C * V + src;
    if (a.C == 4 && a.ld == 4) {
      st4(o, make_float4(__ldg(xs), __ldg(xs + V), __ldg(xs + 2 * V), __ldg(xs + 3 * V)));
    } else {
      for (int c = 0; c < a.C; ++c) o[c] = __ldg(xs + (long long)c * V);
      for (int c = a.C; c < a.ld; ++c) o[c] = 0.f;
    }
    if (a.seg != nullptr) {
      const long long si = (long long)n * V + src;
      const short t = a.seg_bytes == 2 ? __ldg(static_cast<const short*>(a.seg) + si)
                                       : (short)__ldg(static_cast<const signed char*>(a.seg) + si);
      signed char c0, c1, c2;
      if (a.inclusive) {
        c0 = (t == 1 || t == 4); c1 = (t == 1 || t == 2); c2 = (t == 4);   // WT quirk, see below
      } else {
        c0 = (t == 1); c1 = (t == 2); c2 = (t == 4);
      }
      signed char* y = a.yo + (long long)n * 3 * V + v;
      y[0] = c0; y[V] = c1; y[2 * V] = c2;
    }
  }
}

static inline unsigned g1d(long long total) {
  long long b = (total + 255) / 256, cap = (long long)kNumSMs * 16;
  return (unsigned)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace nas3d

using namespace nas3d;

extern "C" {

static int stage_patches_impl(const float* x, const void* seg, int seg_bytes, int N, int C, int D, int H, int W,
                              const int* index_map, int inclusive, float* x_out, int ld_out,
                              signed char* y_out, void* stream) {
  NAS3D_REQUIRE(x && x_out && index_map, "stage_patches: null pointer");
  NAS3D_REQUIRE(N >= 1 && N <= kStageMaxN, "stage_patches: batch %d outside [1,%d]", N, kStageMaxN);
  NAS3D_REQUIRE(C >= 1 && ld_out >= C && D > 0 && H > 0 && W > 0, "stage_patches: bad shape");
  NAS3D_REQUIRE((seg == nullptr) == (y_out == nullptr), "stage_patches: seg and y_out go together");
  const long long V = (long long)D * H * W;
  NAS3D_REQUIRE(V < (1ll << 31), "stage_patches: patch too large");
  StageArgs a;
  a.x = x; a.seg = seg; a.seg_bytes = seg_bytes; a.xo = x_out; a.yo = y_out;
  a.N = N; a.C = C; a.D = D; a.H = H; a.W = W; a.ld = ld_out; a.inclusive = inclusive;
  for (int n = 0; n < N; ++n) {
    const int* m = index_map + 4 * n;
    // the map must stay inside the source patch: check its two extreme corners
    long long lo = m[0], hi = m[0];
    const int ext[3] = {D - 1, H - 1, W - 1};
    for (int k = 0; k < 3; ++k) {
      const long long span = (long long)m[1 + k] * ext[k];
      if (span < 0) lo += span; else hi += span;
    }
    NAS3D_REQUIRE(lo >= 0 && hi < V, "stage_patches: index map of sample %d leaves the patch", n);
    a.base[n] = m[0];
    a.sd[n][0] = m[1]; a.sd[n][1] = m[2]; a.sd[n][2] = m[3];
  }
  NAS3D_REQUIRE(!(C == 4 && ld_out == 4) || aligned16(x_out),
                "stage_patches: x_out must be 16-byte aligned");
  stage_patches_kernel<<<g1d(V * N), 256, 0, (cudaStream_t)stream>>>(a);
  return launched("stage_patches");
}

int nas3d_stage_patches(const float* x, const short* seg, int N, int C, int D, int H, int W,
                        const int* index_map, int inclusive, float* x_out, int ld_out,
                        signed char* y_out, void* stream) {
  return stage_patches_impl(x, seg, 2, N, C, D, H, W, index_map, inclusive, x_out, ld_out, y_out, stream);
}

int nas3d_stage_patches_seg8(const float* x, const signed char* seg, int N, int C, int D, int H, int W,
                             const int* index_map, int inclusive, float* x_out, int ld_out,
                             signed char* y_out, void* stream) {
  return stage_patches_impl(x, seg, 1, N, C, D, H, W, index_map, inclusive, x_out, ld_out, y_out, stream);
}

int nas3d_extract_patches(const float* volume, int C, int D, int H, int W, const int* corners_dev,
                          int B, int Pd, int Ph, int Pw, float* out, int ld_out, void* stream) {
  NAS3D_REQUIRE(C >= 1 && ld_out >= C && B >= 1 && Pd > 0 && Ph > 0 && Pw > 0, "extract_patches: bad shape");
  const long long total = (long long)B * Pd * Ph * Pw;
  extract_patches_kernel<<<g1d(total), 256, 0, (cudaStream_t)stream>>>(volume, C, D, H, W, corners_dev,
                                                                      B, Pd, Ph, Pw, out, ld_out, total);
  return launched("extract_patches");
}

int nas3d_patch_nonzero(const float* volume, int C, int D, int H, int W, const int* corners_dev, int B,
                        int Pd, int Ph, int Pw, int* flags, void* stream) {
  NAS3D_REQUIRE(volume && corners_dev && flags, "patch_nonzero: null pointer");
  NAS3D_REQUIRE(C >= 1 && B >= 1 && B <= 65535 && Pd > 0 && Ph > 0 && Pw > 0, "patch_nonzero: bad shape");
  const long long per_patch = (long long)Pd * Ph * Pw;
  NAS3D_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * B, (cudaStream_t)stream));
  unsigned gx = (unsigned)((per_patch + 255) / 256);
  const unsigned cap = (unsigned)(kNumSMs * 16 / B + 1);
  if (gx > cap) gx = cap;
  patch_nonzero_kernel<<<dim3(gx, B), 256, 0, (cudaStream_t)stream>>>(volume, C, D, H, W, corners_dev,
                                                                      Pd, Ph, Pw, flags, per_patch);
  return launched("patch_nonzero");
}

int nas3d_stitch_labels(const float* const* patch_preds_dev, const int* patch_nonzero_dev, int ld_pred,
                        const int* corners_dev, int B, int Pd, int Ph,
                        int Pw, int Db, int Hb, int Wb, int D, int H, int W, int off_d, int off_h,
                        int off_w, float threshold, int inclusive, const unsigned char* skull_mask,
                        unsigned char* labels, double* stitched, void* stream) {
  NAS3D_REQUIRE(patch_preds_dev && corners_dev && labels, "stitch_labels: null pointer");
  NAS3D_REQUIRE(ld_pred >= 3 && B >= 1, "stitch_labels: predictions need 3 channels");
  NAS3D_REQUIRE(off_d >= 0 && off_h >= 0 && off_w >= 0 && off_d + Db <= D && off_h + Hb <= H &&
                    off_w + Wb <= W, "stitch_labels: brain box outside the volume");
  StitchArgs A{patch_preds_dev, patch_nonzero_dev, corners_dev, B, Pd, Ph, Pw, ld_pred, Db, Hb, Wb, D, H, W, off_d, off_h, off_w,
               threshold, inclusive, skull_mask, labels, stitched};
  const long long total = (long long)D * H * W;
  stitch_labels_kernel<<<g1d(total), 256, 0, (cudaStream_t)stream>>>(A, total);
  return launched("stitch_labels");
}

int nas3d_seg_to_masks(const short* seg, int N, long long V, int inclusive, float* masks, void* stream) {
  const long long total = (long long)N * V;
  seg_to_masks_kernel<<<g1d(total), 256, 0, (cudaStream_t)stream>>>(seg, V, total, inclusive, masks);
  return launched("seg_to_masks");
}

}  // extern "C"
