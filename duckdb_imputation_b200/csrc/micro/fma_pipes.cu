// fma_pipes.cu -- microbenchmark: FMA throughput per SM of FFMA, FFMA2 (fma.rn.f32x2) and a mix,
// at 2 / 4 / 8 warps per SMSP.  Answers whether packed FFMA2 and scalar FFMA share one pipe.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ffma2(unsigned long long &d, unsigned long long a, unsigned long long b) {
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ void ffma(float &d, float a, float b) { asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(d) : "f"(a), "f"(b)); }

template <int MODE>
__global__ void k(float *out, int iters, float a, float b) {
  unsigned long long A = ((unsigned long long)__float_as_uint(a) << 32) | __float_as_uint(a);
  unsigned long long B = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
  unsigned long long p[16];
  float s[16];
#pragma unroll
  for (int i = 0; i < 16; i++) { p[i] = i; s[i] = (float)i; }
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) {  // 32 scalar FMA
#pragma unroll
      for (int i = 0; i < 16; i++) ffma(s[i], a, b);
#pragma unroll
      for (int i = 0; i < 16; i++) ffma(s[i], b, a);
    } else if (MODE == 1) {  // 16 packed = 32 FMA
#pragma unroll
      for (int i = 0; i < 16; i++) ffma2(p[i], A, B);
    } else {  // 8 packed + 16 scalar interleaved = 32 FMA
#pragma unroll
      for (int i = 0; i < 8; i++) { ffma2(p[i], A, B); ffma(s[2 * i], a, b); ffma(s[2 * i + 1], b, a); }
    }
  }
  float r = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) r += s[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
  if (r == 12345.678f) out[0] = r;
}
int main() {
  float *out; cudaMalloc(&out, 4);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 200000;
  const char *names[3] = {"FFMA x32", "FFMA2 x16", "FFMA2 x8 + FFMA x16"};
  for (int warps : {4, 8, 16, 32})
    for (int mode = 0; mode < 3; mode++) {
      for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<sms, warps * 32>>>(out, iters, 1.0001f, 0.9999f);
        if (mode == 1) k<1><<<sms, warps * 32>>>(out, iters, 1.0001f, 0.9999f);
        if (mode == 2) k<2><<<sms, warps * 32>>>(out, iters, 1.0001f, 0.9999f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma = (double)iters * 32 * warps * 32 * sms;
      printf("warps/SM=%2d %-22s %8.3f ms  %7.2f TFMA/s  %6.1f FMA/clk/SM (at %d MHz nominal)\n", warps, names[mode], ms, fma / ms / 1e9,
             fma / ms / 1e3 / sms / (clk / 1e3) / 1e3, clk / 1000);
    }
  return 0;
}
