// tc_probe.cu -- probe of tcgen05.mma kind::tf32 with hand-built descriptors: D[64x24] = A[64xK] * B[24xK]^T,
// K-major no-swizzle ("interleave") canonical layout, cta_group::1.  Prints how D rows map to TMEM lanes.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__host__ __device__ inline uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
  return d;                // layout_type = 0 (no swizzle), base_offset = 0, lbo_mode = 0
}

constexpr int M = 64, N = 24, K = 32;
constexpr uint32_t LBO = 128, SBO = (K / 4) * 128;

__global__ void probe(const float *A, const float *B, float *out, int swap_lbo_sbo) {
  __shared__ __align__(1024) unsigned char sa[M / 8 * SBO];
  __shared__ __align__(1024) unsigned char sb[N / 8 * SBO];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < M * K; i += blockDim.x) {
    const int m = i / K, k = i % K;
    *(float *)(sa + (m / 8) * SBO + (k / 4) * LBO + (m % 8) * 16 + (k % 4) * 4) = A[i];
  }
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
    const int n = i / K, k = i % K;
    *(float *)(sb + (n / 8) * SBO + (k / 4) * LBO + (n % 8) * 16 + (k % 4) * 4) = B[i];
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t taddr = tmem_base;
  if (threadIdx.x == 32) {  // one thread issues
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    for (int ks = 0; ks < K / 8; ks++) {
      const uint32_t lbo = swap_lbo_sbo ? SBO : LBO, sbo = swap_lbo_sbo ? LBO : SBO;
      const uint64_t ad = make_desc(smem_u32(sa) + ks * 2 * LBO, lbo, sbo);
      const uint64_t bd = make_desc(smem_u32(sb) + ks * 2 * LBO, lbo, sbo);
      const uint32_t acc = ks > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(taddr),
          "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // everyone waits for the MMAs
  {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok)
                   : "r"(smem_u32(&bar))
                   : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 4) {
    for (int c0 = 0; c0 < N; c0 += 8) {
      uint32_t r[8];
      const uint32_t addr = taddr + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(addr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; j++) out[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(taddr) : "memory");
}

int main() {
  std::vector<float> A(M * K), B(N * K), D(M * N, 0.f), out(128 * N);
  for (int m = 0; m < M; m++)
    for (int k = 0; k < K; k++) A[m * K + k] = (float)((m * 3 + k * 5) % 7 - 3);
  for (int n = 0; n < N; n++)
    for (int k = 0; k < K; k++) B[n * K + k] = (float)((n * 2 + k) % 5 - 2);
  for (int m = 0; m < M; m++)
    for (int n = 0; n < N; n++)
      for (int k = 0; k < K; k++) D[m * N + n] += A[m * K + k] * B[n * K + k];
  float *dA, *dB, *dO;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, out.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  for (int swap = 0; swap < 2; swap++) {
    cudaMemset(dO, 0, out.size() * 4);
    probe<<<1, 128>>>(dA, dB, dO, swap);
    cudaError_t e = cudaDeviceSynchronize();
    printf("swap_lbo_sbo=%d: %s\n", swap, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
    // which D row does each TMEM lane hold?
    int matched = 0;
    for (int l = 0; l < 128; l++) {
      int row = -1;
      for (int m = 0; m < M && row < 0; m++) {
        bool eq = true;
        for (int n = 0; n < N; n++) eq = eq && out[l * N + n] == D[m * N + n];
        if (eq) row = m;
      }
      if (row >= 0) matched++;
      if (l < 20 || l % 16 == 0 || row >= 0) printf("lane %3d -> D row %2d   (first vals %.0f %.0f %.0f | expect row0 %.0f %.0f %.0f)\n", l, row, out[l * N], out[l * N + 1], out[l * N + 2], D[0], D[1], D[2]);
    }
    printf("lanes matching some D row: %d\n", matched);
  }
  return 0;
}
