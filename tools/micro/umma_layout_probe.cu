// Probe of tcgen05.mma kind::tf32 operand layouts on sm_100a: ONE MMA (M=128, N=32, K=8) with the
// shared-memory images of A and B built on the host under a layout hypothesis; D read back from
// TMEM and compared with the host product.  Settles which (major, swizzle, LBO, SBO) combinations
// the hardware accepts for 32-bit operands (the PTX documentation is not available offline).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/micro/umma_layout_probe tools/micro/umma_layout_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Probe {
  uint32_t a_lbo, a_sbo, a_layout;   // layout: 0 none, 2 SW128, 4 SW64, 6 SW32
  uint32_t b_lbo, b_sbo, b_layout;
  uint32_t idesc;
  uint32_t a_adv, b_adv;             // byte advance of the start address (K block inside a swizzle atom)
};

__global__ void __launch_bounds__(128) probe_kernel(const float* imgA, const float* imgB, float* D, Probe P, int N) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* sa = reinterpret_cast<float*>(smem);               // 16 KB
  float* sb = reinterpret_cast<float*>(smem + 16384);       // 16 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 32768 + 16);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 4096; i += 128) { sa[i] = imgA[i]; sb[i] = imgB[i]; }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;\n" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tb = *slot;
  if (tid == 0) {
    auto desc = [](uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
      return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
             ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
    };
    const uint64_t da = desc(smem_u32(sa) + P.a_adv, P.a_lbo, P.a_sbo, P.a_layout);
    const uint64_t db = desc(smem_u32(sb) + P.b_adv, P.b_lbo, P.b_sbo, P.b_layout);
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tb), "l"(da), "l"(db),
        "r"(P.idesc), "r"(0u)
        : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
                 : "memory");
  }
  asm volatile(
      "{\n\t.reg .pred p;\n\tPW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra PD;\n\tbra PW;\n\tPD:\n\t}\n" ::"r"(
          smem_u32(bar))
      : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  for (int c8 = 0; c8 < N / 8; ++c8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(tb + ((uint32_t)(warp * 32) << 16) + c8 * 8));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    for (int i = 0; i < 8; ++i) D[tid * N + c8 * 8 + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;\n" ::"r"(tb) : "memory");
}

static float tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

// byte offset of element (r, k) of an R x 8 operand under a layout hypothesis
static uint32_t off_k_none(int r, int k) { return (r / 8) * 256 + (k / 4) * 128 + (r % 8) * 16 + (k % 4) * 4; }   // LBO 128, SBO 256
static uint32_t off_mn_none(int r, int k) { return (r / 4) * 128 + k * 16 + (r % 4) * 4; }                          // SBO 128
static uint32_t off_mn_sw128(int r, int k) { return (r / 32) * 1024 + k * 128 + ((((r % 32) / 4) ^ (k % 8)) * 16) + (r % 4) * 4; }   // LBO 1024
static uint32_t off_k_sw32(int r, int k) { return (r / 8) * 256 + (r % 8) * 32 + ((((k / 4)) ^ ((r % 8) / 4 % 2)) * 16) + (k % 4) * 4; }

// K-major SWIZZLE_128B: row r = 128 bytes = 32 k, 16-byte chunk index XOR (r % 8); 8-row atoms of 1024 B
static uint32_t off_k_sw128(int r, int k) { return (r / 8) * 1024 + (r % 8) * 128 + ((((k / 4)) ^ (r % 8)) * 16) + (k % 4) * 4; }

int main() {
  const int M = 128, N = 32, KF = 32;
  std::vector<float> A(M * KF), B(N * KF);
  for (int i = 0; i < M * KF; ++i) A[i] = tf32((float)((i * 37 % 101) - 50) / 16.f);
  for (int i = 0; i < N * KF; ++i) B[i] = tf32((float)((i * 53 % 97) - 48) / 8.f);
  float *dA, *dB, *dD;
  cudaMalloc(&dA, 16384); cudaMalloc(&dB, 16384); cudaMalloc(&dD, M * N * 4);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 33792);
  struct Case { const char* name; int a_kind, b_kind, kb; Probe P; };
  const uint32_t base = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  // kinds: 0 K-major none (k block kb lives at +kb*256 per 8-row group: image holds only that block), 3 K-major SW128
  std::vector<Case> cases = {
      {"A K/none  B K/none  kb0               ", 0, 0, 0, {128, 256, 0, 128, 256, 0, base, 0, 0}},
      {"A K/sw128 B K/sw128 kb0 lbo0 sbo1024  ", 3, 3, 0, {0, 1024, 2, 0, 1024, 2, base, 0, 0}},
      {"A K/sw128 B K/sw128 kb1 (+32B)        ", 3, 3, 1, {0, 1024, 2, 0, 1024, 2, base, 32, 32}},
      {"A K/sw128 B K/sw128 kb3 (+96B)        ", 3, 3, 3, {0, 1024, 2, 0, 1024, 2, base, 96, 96}},
      {"A K/sw128 B K/sw128 kb2 lbo16         ", 3, 3, 2, {16, 1024, 2, 16, 1024, 2, base, 64, 64}},
      {"A K/sw128 B K/none  kb1               ", 3, 0, 1, {0, 1024, 2, 128, 256, 0, base, 32, 0}},
      {"A K/none  B K/sw128 kb2               ", 0, 3, 2, {128, 256, 0, 0, 1024, 2, base, 0, 64}},
  };
  for (auto& c : cases) {
    std::vector<float> ia(4096, 0.f), ib(4096, 0.f), out(M * N, -1.f), ref(M * N);
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        float s = 0; for (int k = 0; k < 8; ++k) s += A[m * KF + c.kb * 8 + k] * B[n * KF + c.kb * 8 + k];
        ref[m * N + n] = s;
      }
    for (int m = 0; m < M; ++m) for (int k = 0; k < KF; ++k) {
      if (c.a_kind == 0) { if (k / 8 == c.kb) ia[off_k_none(m, k % 8) / 4] = A[m * KF + k]; }
      else ia[off_k_sw128(m, k) / 4] = A[m * KF + k];
    }
    for (int n = 0; n < N; ++n) for (int k = 0; k < KF; ++k) {
      if (c.b_kind == 0) { if (k / 8 == c.kb) ib[off_k_none(n, k % 8) / 4] = B[n * KF + k]; }
      else ib[off_k_sw128(n, k) / 4] = B[n * KF + k];
    }
    cudaMemcpy(dA, ia.data(), 16384, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, ib.data(), 16384, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xFF, M * N * 4);
    probe_kernel<<<1, 128, 33792>>>(dA, dB, dD, c.P, N);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", c.name, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(out.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0; int nz = 0;
    for (int i = 0; i < M * N; ++i) {
      maxerr = fmax(maxerr, fabs((double)out[i] - ref[i])); maxref = fmax(maxref, fabs((double)ref[i]));
      nz += out[i] != 0.f;
    }
    printf("%s: max|D-ref| %.4g of %.4g  nonzero %d/%d  D[0][0..2] %g %g %g (ref %g %g %g)  %s\n", c.name, maxerr, maxref,
           nz, M * N, out[0], out[1], out[2], ref[0], ref[1], ref[2], maxerr <= 1e-3 * maxref ? "MATCH" : "-");
  }
  printf("(MN-major operands - instruction descriptor bits 15 / 16 - were probed first: every combination of\n"
         " no-swizzle / 128B-swizzle and LBO / SBO order returned D == 0 for kind::tf32: 32-bit operands are K-major only)\n");
  return 0;
}
