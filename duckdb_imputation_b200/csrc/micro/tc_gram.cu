// tc_gram.cu -- MEASUREMENT: the Gram matrix X^T X of n = 20 FLOAT columns on the 5th-generation tensor cores
// (tcgen05.mma kind::tf32, accumulator in TMEM), inside the same bulk-copy ring as gram_scan_kernel, to settle with a
// number whether the dense contraction of sum_to_triple_20_0 belongs on tcgen05.
//
// G = X^T X with K = rows as the reduction: A = X^T (M x K, K contiguous: a column of the table IS a K-major row) and
// B = X^T (N x K) -- the SoA columns are the operands, no transpose of the table.  fp32 inputs need <= 1e-5 relative:
// x = hi + lo with hi = tf32(x) (low 13 mantissa bits cleared), lo = x - hi, and
//     G = hi.hi + lo.hi + (lo.hi)^T  (+ lo.lo ~ 2^-22, dropped)
// as ONE instruction shape per 8 rows:  D[64 x 24] += [H (32 rows: 20 + padding) ; L (32 rows)][64 x 8] * H[24 x 8]^T.
// Per tile of KT rows:  producer warp: 20 bulk async copies (cp.async.bulk) into a 3-stage raw ring;
//   4 convert warps: raw fp32 -> hi / lo in the UMMA canonical K-major no-swizzle layout (8-row x 16-byte core
//   matrices; the raw SoA chunks do not have that layout, and the split needs a CUDA-core pass anyway);
//   one thread: KT / 8 tcgen05.mma, commit -> the operand buffer is free again (double-buffered).
// Epilogue: tcgen05.ld of the 64 x 24 accumulator, per-CTA partials summed on the host in fp64.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_gram tc_gram.cu ;  ./tc_gram [rows_resident] [passes] [scans]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("%s failed: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__);   \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

constexpr int NCOL = 20, KT = 256, STAGES = 3, PITCH = KT + 4;  // raw column pitch in floats: +16 bytes (bank skew)
constexpr int M = 64, N = 24;
constexpr uint32_t LBO = 128, SBO = (KT / 4) * 128;             // K-chunk stride, 8-row group stride (bytes)
constexpr int OPS_BYTES = 8 * (KT / 4) * 128;                   // 8 groups of 8 rows: H 0-3, L 4-7
constexpr int RAW_BYTES = NCOL * PITCH * 4;
constexpr int THREADS = 192;                                    // warp 0 producer, warp 1 MMA, warps 2-5 convert / epilogue

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46);  // version 1 (Blackwell), no swizzle
}

struct Cols {
  const float *c[NCOL];
};

__global__ void __launch_bounds__(THREADS, 1) tc_gram_kernel(Cols cols, unsigned long long tiles_resident, unsigned long long n_tiles,
                                                             float *out /* [grid][64][24] */) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *ops = smem;                                           // [2][OPS_BYTES]
  float *raw = reinterpret_cast<float *>(smem + 2 * OPS_BYTES);         // [STAGES][NCOL][PITCH]
  __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES], ops_ready[2], ops_free[2], done;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 4);
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(&ops_ready[b], 4);
      mbar_init(&ops_free[b], 1);
    }
    mbar_init(&done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 2 * OPS_BYTES / 16; i += THREADS) reinterpret_cast<uint4 *>(ops)[i] = make_uint4(0, 0, 0, 0);  // padding rows
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t taddr = tmem_base;
  // this CTA's tiles: t = blockIdx.x, + gridDim.x, ...
  const unsigned long long my_tiles = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == 0) {  // ---- producer: lane c issues column c
    for (unsigned long long i = 0; i < my_tiles; i++) {
      const int s = (int)(i % STAGES);
      if (i >= STAGES) mbar_wait(&empty[s], (uint32_t)((i / STAGES - 1) & 1));
      if (lane == 0) mbar_expect(&full[s], NCOL * KT * 4);
      __syncwarp();
      const unsigned long long t = (blockIdx.x + i * gridDim.x) % tiles_resident;
      if (lane < NCOL) bulk_g2s(raw + ((size_t)s * NCOL + lane) * PITCH, cols.c[lane] + t * KT, KT * 4, &full[s]);
    }
  } else if (warp == 1) {  // ---- MMA issuer
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
      for (unsigned long long i = 0; i < my_tiles; i++) {
        const int b = (int)(i & 1);
        mbar_wait(&ops_ready[b], (uint32_t)((i >> 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t base = smem_u32(ops + (size_t)b * OPS_BYTES);
#pragma unroll 4
        for (int ks = 0; ks < KT / 8; ks++) {
          const uint64_t ad = make_desc(base + ks * 2 * LBO, LBO, SBO), bd = ad;  // B = the H rows (groups 0-2) of the same buffer
          const uint32_t acc = (i > 0 || ks > 0) ? 1u : 0u;
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(taddr),
                       "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&ops_free[b])) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
    }
  } else {  // ---- convert warps 2..5: raw fp32 -> hi / lo operands
    const int cw = warp - 2, m8 = lane & 7, kq = lane >> 3;  // lane = (row within an 8-row group, one of 4 K chunks)
    for (unsigned long long i = 0; i < my_tiles; i++) {
      const int s = (int)(i % STAGES), b = (int)(i & 1);
      mbar_wait(&full[s], (uint32_t)((i / STAGES) & 1));
      if (i >= 2) mbar_wait(&ops_free[b], (uint32_t)(((i >> 1) - 1) & 1));
      const float *rw = raw + (size_t)s * NCOL * PITCH;
      unsigned char *ob = ops + (size_t)b * OPS_BYTES;
      // K chunks (4 rows each): KT / 4 = 64; this warp takes chunks cw * 16 .. cw * 16 + 15, four at a time
      for (int g = 0; g < 3; g++) {
        const int m = g * 8 + m8;
        if (m < NCOL) {
#pragma unroll
          for (int k4 = 0; k4 < 4; k4++) {
            const int kc = cw * 16 + k4 * 4 + kq;
            const float4 x = *reinterpret_cast<const float4 *>(rw + (size_t)m * PITCH + 4 * kc);
            float4 h, l;
            h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u), l.x = x.x - h.x;
            h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u), l.y = x.y - h.y;
            h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u), l.z = x.z - h.z;
            h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u), l.w = x.w - h.w;
            *reinterpret_cast<float4 *>(ob + (size_t)g * SBO + (size_t)kc * LBO + m8 * 16) = h;
            *reinterpret_cast<float4 *>(ob + (size_t)(4 + g) * SBO + (size_t)kc * LBO + m8 * 16) = l;
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&ops_ready[b]);
        mbar_arrive(&empty[s]);
      }
    }
    // ---- epilogue: this warp's TMEM sub-partition (warp id % 4): D row 16 * sp + l sits in lane 32 * sp + l, l < 16
    mbar_wait(&done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int sp = warp & 3;
    for (int c0 = 0; c0 < N; c0 += 8) {
      uint32_t r[8];
      const uint32_t addr = taddr + ((uint32_t)(sp * 32) << 16) + c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (lane < 16 && my_tiles)
        for (int j = 0; j < 8; j++) out[((size_t)blockIdx.x * M + sp * 16 + lane) * N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(taddr) : "memory");
}

__global__ void fill_kernel(float *p, size_t n, unsigned long long seed) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned long long z = seed * 0xD1342543DE82EF95ull + i + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    p[i] = (float)(z >> 40) * (1.0f / 16777216.0f);
  }
}

int main(int argc, char **argv) {
  const size_t rows_resident = (argc > 1 ? (size_t)atoll(argv[1]) : (size_t)250'000'000) / KT * KT;
  const int passes = argc > 2 ? atoi(argv[2]) : 4, scans = argc > 3 ? atoi(argv[3]) : 10;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  Cols cols;
  for (int c = 0; c < NCOL; c++) {
    float *p;
    CK(cudaMalloc(&p, rows_resident * 4));
    fill_kernel<<<sms * 8, 256>>>(p, rows_resident, 1000 + c);
    cols.c[c] = p;
  }
  CK(cudaDeviceSynchronize());
  const size_t smem = 2 * OPS_BYTES + (size_t)STAGES * RAW_BYTES;
  CK(cudaFuncSetAttribute(tc_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  float *d_out;
  CK(cudaMalloc(&d_out, (size_t)sms * M * N * 4));
  const unsigned long long tiles_resident = rows_resident / KT;

  // ---- correctness over prefixes of growing length: the accumulator lives in TMEM as fp32 for the whole scan of a CTA,
  //      so the error grows with the number of rows a CTA adds up before its (single) epilogue
  for (size_t tiles_per_cta : {(size_t)1, (size_t)8, (size_t)64, (size_t)256}) {
    const size_t check_rows = std::min<size_t>(rows_resident, tiles_per_cta * sms * KT) / KT * KT;
    CK(cudaMemset(d_out, 0, (size_t)sms * M * N * 4));
    tc_gram_kernel<<<sms, THREADS, smem>>>(cols, tiles_resident, check_rows / KT, d_out);
    CK(cudaDeviceSynchronize());
    std::vector<float> out((size_t)sms * M * N);
    CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
    std::vector<double> D((size_t)M * N, 0.0);
    for (int b = 0; b < sms; b++)
      for (int i = 0; i < M * N; i++) D[i] += out[(size_t)b * M * N + i];
    std::vector<std::vector<float>> h(NCOL, std::vector<float>(check_rows));
    for (int c = 0; c < NCOL; c++) CK(cudaMemcpy(h[c].data(), cols.c[c], check_rows * 4, cudaMemcpyDeviceToHost));
    double worst = 0.0, worst_hh = 0.0;
    for (int i = 0; i < NCOL; i++)
      for (int j = i; j < NCOL; j++) {
        double ref = 0.0;
        for (size_t r = 0; r < check_rows; r++) ref += (double)h[i][r] * (double)h[j][r];
        const double got = D[i * N + j] + D[(32 + i) * N + j] + D[(32 + j) * N + i];
        worst = std::max(worst, std::abs(got - ref) / std::abs(ref));
        worst_hh = std::max(worst_hh, std::abs(D[i * N + j] - ref) / std::abs(ref));
      }
    printf("tc_gram check: %9zu rows (%4zu rows per CTA between epilogues): max rel. error vs fp64 = %.3e  (hi.hi term alone: %.3e)\n",
           check_rows, tiles_per_cta * KT, worst, worst_hh);
  }
  // ---- timing: `scans` back-to-back scans of passes x rows_resident rows
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const unsigned long long n_tiles = tiles_resident * passes;
  for (int w = 0; w < 2; w++) tc_gram_kernel<<<sms, THREADS, smem>>>(cols, tiles_resident, n_tiles, d_out);
  CK(cudaDeviceSynchronize());
  for (int k = 0; k < scans; k++) {
    CK(cudaEventRecord(e0));
    tc_gram_kernel<<<sms, THREADS, smem>>>(cols, tiles_resident, n_tiles, d_out);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double rows = (double)n_tiles * KT;
    printf("tc_gram scan %2d: %.0f rows  %8.3f ms  %7.2f G rows/s  %7.1f GB/s\n", k, rows, ms, rows / ms / 1e6, rows * NCOL * 4 / ms / 1e6);
  }
  return 0;
}
