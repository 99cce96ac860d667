// slab_launch.h -- host-side launch interface of slab_scan_kernel (slab_kernels.cuh); the
// instantiations (n = 0..32, triple / NB) live in slab_inst.cu translation units.
#pragma once
#include <cuda_runtime.h>

#include "pair_hash.cuh"
#include "state_layout.h"

namespace cfb {

struct SlabLaunchParams {
  ScanCols cols;
  const Layout *d_lay;
  unsigned long long rows;
  int do_numeric;
  int flush_tiles;
  float *slab;
  double *f64;
  unsigned long long *u64;
  int *err;
  int grid;
  cudaStream_t stream;
  PairHash hash;
};

template <int N, int KIND>
cudaError_t slab_launch(const SlabLaunchParams &p);

}  // namespace cfb
