// stream_ceiling.cu -- microbenchmark: how fast can one persistent CTA per SM pull `ncols`
// column streams from HBM with 1-D bulk async copies (the feed of gram_scan_kernel), with
// the consumer doing nothing but releasing stages?  And the same bytes with plain LDG.128.
// Prints GB/s per configuration.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../ptx_sm100.cuh"
using namespace cfb;

struct Cols { const float *p[32]; };

__global__ void __launch_bounds__(288, 1) bulk_ring(Cols cols, int ncols, unsigned long long rows, int tr, int stages,
                                                    int evict_first, float *sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  float *ring = (float *)smem;
  uint64_t *full = (uint64_t *)(ring + (size_t)stages * ncols * tr);
  uint64_t *empty = full + stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; s++) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 8); }
    ptx::fence_mbar_init();
  }
  __syncthreads();
  const unsigned long long tiles = rows / tr;
  float acc = 0.f;
  if (warp == 8) {
    if (lane == 0) {
      uint64_t pol = evict_first ? ptx::policy_evict_first() : 0;
      int stage = 0; uint32_t phase = 0;
      for (unsigned long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        ptx::mbar_wait(&empty[stage], phase ^ 1u);
        ptx::mbar_arrive_expect_tx(&full[stage], (uint32_t)tr * 4u * ncols);
        float *dst = ring + (size_t)stage * ncols * tr;
        for (int c = 0; c < ncols; c++) {
          if (evict_first) ptx::bulk_g2s(dst + c * tr, cols.p[c] + t * tr, tr * 4u, &full[stage], pol);
          else asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                            ::"r"(ptx::smem_u32(dst + c * tr)), "l"(cols.p[c] + t * tr), "r"(tr * 4u), "r"(ptx::smem_u32(&full[stage])) : "memory");
        }
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    int stage = 0; uint32_t phase = 0;
    for (unsigned long long t = blockIdx.x; t < tiles; t += gridDim.x) {
      ptx::mbar_wait(&full[stage], phase);
      acc += ring[(size_t)stage * ncols * tr + threadIdx.x];
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&empty[stage]);
      if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
  }
  if (acc == 123.456f) sink[0] = acc;
}

__global__ void __launch_bounds__(512) ldg_stream(Cols cols, int ncols, unsigned long long rows, float *sink) {
  float acc = 0.f;
  const unsigned long long n4 = rows / 4;
  for (int c = 0; c < ncols; c++) {
    const float4 *p = (const float4 *)cols.p[c];
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (unsigned long long)gridDim.x * blockDim.x * 4) {
      float4 a = __ldcs(p + i);
      float4 b = (i + (unsigned long long)gridDim.x * blockDim.x < n4) ? __ldcs(p + i + (unsigned long long)gridDim.x * blockDim.x) : make_float4(0, 0, 0, 0);
      float4 c2 = (i + 2ull * gridDim.x * blockDim.x < n4) ? __ldcs(p + i + 2ull * gridDim.x * blockDim.x) : make_float4(0, 0, 0, 0);
      float4 d = (i + 3ull * gridDim.x * blockDim.x < n4) ? __ldcs(p + i + 3ull * gridDim.x * blockDim.x) : make_float4(0, 0, 0, 0);
      acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w + c2.x + c2.y + c2.z + c2.w + d.x + d.y + d.z + d.w;
    }
  }
  if (acc == 123.456f) sink[0] = acc;
}

int main(int argc, char **argv) {
  const unsigned long long rows = argc > 1 ? atoll(argv[1]) : 200000000ull;
  const int ncols = 20;
  Cols cols{};
  for (int c = 0; c < ncols; c++) { cudaMalloc((void **)&cols.p[c], rows * 4); cudaMemset((void *)cols.p[c], 0, rows * 4); }
  float *sink; cudaMalloc(&sink, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaFuncSetAttribute(bulk_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
  const double gb = (double)rows * 4 * ncols / 1e9;
  auto time = [&](auto launch) { launch(); cudaDeviceSynchronize(); float best = 1e9; for (int i = 0; i < 3; i++) { cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best; } return best; };
  for (int ef = 0; ef < 2; ef++)
    for (int tr : {256, 512, 1024, 2048})
      for (int stages : {2, 3, 4, 5, 8}) {
        size_t smem = (size_t)stages * ncols * tr * 4 + 2 * stages * 8;
        if (smem > 226 * 1024) continue;
        for (int grid : {sms, 2 * sms}) {
          if (grid == 2 * sms && smem > 110 * 1024) continue;
          float ms = time([&] { bulk_ring<<<grid, 288, smem>>>(cols, ncols, rows, tr, stages, ef, sink); });
          cudaError_t e = cudaGetLastError();
          printf("bulk evict_first=%d tr=%4d (%5d B/col) stages=%d grid=%3d smem=%6zu : %7.3f ms %7.1f GB/s %s\n", ef, tr, tr * 4, stages, grid, smem, ms, gb / ms * 1e3, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
      }
  for (int grid : {sms * 2, sms * 4, sms * 8}) {
    float ms = time([&] { ldg_stream<<<grid, 512>>>(cols, ncols, rows, sink); });
    printf("ldg.128 x4 grid=%4d : %7.3f ms %7.1f GB/s\n", grid, ms, gb / ms * 1e3);
  }
  // sustained: 12 back-to-back launches of the best feed configurations (power-capped regime)
  for (int tr : {768, 1024}) {
    const int stages = tr == 768 ? 3 : 2;
    const size_t smem = (size_t)stages * ncols * tr * 4 + 2 * stages * 8;
    bulk_ring<<<sms, 288, smem>>>(cols, ncols, rows, tr, stages, 0, sink);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < 12; i++) bulk_ring<<<sms, 288, smem>>>(cols, ncols, rows, tr, stages, 0, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms12;
    cudaEventElapsedTime(&ms12, e0, e1);
    printf("sustained x12: bulk tr=%d stages=%d (single issuing thread): %7.3f ms per launch %7.1f GB/s\n", tr, stages, ms12 / 12, gb / (ms12 / 12) * 1e3);
  }
  // copy baseline (what MEASURED_PEAKS counts: read+write bytes)
  float ms = time([&] { cudaMemcpyAsync((void *)cols.p[1], cols.p[0], rows * 4, cudaMemcpyDeviceToDevice); });
  printf("cudaMemcpy D2D %.3f ms -> %.1f GB/s (read+write)\n", ms, 2.0 * rows * 4 / 1e9 / ms * 1e3);
  return 0;
}
