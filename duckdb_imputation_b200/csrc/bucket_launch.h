// bucket_launch.h -- host-side launch interface of bucket_sum_kernel (bucket_kernels.cuh); the
// instantiations (n = 0..32) live in bucket_inst.cu translation units.
#pragma once
#include <cuda_runtime.h>

#include "state_layout.h"

namespace cfb {

struct BucketLaunchParams {
  ScanCols cols;
  const Layout *lay;  // host copy
  unsigned long long rows;
  int tile_rows, fold_tiles, grid;
  int smem_max;       // cudaDevAttrMaxSharedMemoryPerBlockOptin - 1 KB
  size_t smem_bytes;
  float *slab;
  double *f64;
  unsigned long long *u64;
  int *err;
  cudaStream_t stream;
};

template <int N>
cudaError_t bucket_launch(const BucketLaunchParams &p);

}  // namespace cfb
