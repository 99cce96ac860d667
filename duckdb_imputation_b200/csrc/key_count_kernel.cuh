// key_count_kernel.cuh -- categorical part of the Naive-Bayes ring (included by cofactor_b200.cu only).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "bucket_kernels.cuh"
#include "state_layout.h"

namespace cfb {

// Key counts only (the Naive-Bayes ring: sum_to_nb_agg.cpp:124-145 keeps `map[key][0] += 1` per categorical
// column and nothing else): a shared-memory histogram over (slot, column, key), one integer ATOMS per row and
// column, folded into the u64 state once at the end of the CTA's rows (a CTA sees < 2^32 rows).
struct KeyCountArgs {
  ScanCols cols;
  unsigned long long n_rows;
  int m, total_dom, n_groups;
  long long U;
  int lo[kMaxCat], dom[kMaxCat], cat_off[kMaxCat + 1];
  unsigned long long *u64;
  int *err;
};

__global__ void __launch_bounds__(kBucketThreads, 1) key_count_kernel(const __grid_constant__ KeyCountArgs a) {
  extern __shared__ unsigned key_hist[];  // [n_groups * total_dom]
  const int D = a.n_groups * a.total_dom, m = a.m;
  for (int i = threadIdx.x; i < D; i += kBucketThreads) key_hist[i] = 0;
  __syncthreads();
  bool bad = false;
  for (unsigned long long r = (unsigned long long)blockIdx.x * kBucketThreads + threadIdx.x; r < a.n_rows;
       r += (unsigned long long)gridDim.x * kBucketThreads) {
    int gbase = 0;
    if (a.cols.group) {
      const int g = a.cols.group[r];
      if (g < 0 || g >= a.n_groups) {  // < 0: filtered row
        if (g > 0) atomicExch(a.err, 2);
        continue;
      }
      gbase = g * a.total_dom;
    }
    for (int c0 = 0; c0 < m; c0 += 4) {  // keys four columns at a time: the loads are independent
      unsigned s[4];
#pragma unroll
      for (int e = 0; e < 4; e++) s[e] = c0 + e < m ? (unsigned)(a.cols.cat[c0 + e][r] - a.lo[c0 + e]) : 0u;
#pragma unroll
      for (int e = 0; e < 4; e++)
        if (c0 + e < m) {
          if (s[e] < (unsigned)a.dom[c0 + e]) atomicAdd(&key_hist[gbase + a.cat_off[c0 + e] + s[e]], 1u);
          else bad = true;
        }
    }
  }
  if (bad) atomicExch(a.err, 1);
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += kBucketThreads) {
    const unsigned v = key_hist[i];
    if (v) red_u64(a.u64 + (i / a.total_dom) * a.U + 1 + i % a.total_dom, v);
  }
}

}  // namespace cfb
