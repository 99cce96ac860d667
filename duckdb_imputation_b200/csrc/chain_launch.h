// chain_launch.h -- host-side launch interface of chain_sum_kernel (chain_kernels.cuh); the instantiations
// (n = 0..32) live in chain_inst.cu translation units.
#pragma once
#include <cuda_runtime.h>

#include "state_layout.h"

namespace cfb {

struct ChainLaunchParams {
  ScanCols cols;
  const Layout *lay;  // host copy
  unsigned long long rows;
  int tile_rows, fold_tiles, sub_shift, head_cap, adaptive, grid;
  int smem_max;  // cudaDevAttrMaxSharedMemoryPerBlockOptin - 1 KB
  size_t smem_bytes;
  float *slab;
  unsigned *cnt_slab;
  double *f64;
  unsigned long long *u64;
  int *err;
  unsigned char *packed;  // nullable
  unsigned long long packed_stride;
  cudaStream_t stream;
};

template <int N>
cudaError_t chain_launch(const ChainLaunchParams &p);

}  // namespace cfb
