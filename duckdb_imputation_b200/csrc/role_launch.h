// role_launch.h -- host-side launch interface of role_scan_kernel (role_kernels.cuh); the
// instantiations (n = 0..32, 16- / 32-bit cells) live in role_inst.cu translation units.
#pragma once
#include <cuda_runtime.h>

#include "state_layout.h"

namespace cfb {

struct RolePlan;

struct RoleLaunchParams {
  ScanCols cols;
  const Layout *lay;     // host copy
  const RolePlan *plan;  // host copy (travels in the kernel parameters)
  unsigned long long rows;
  int chunk_rows, pair_fold_chunks, n_roles, n_reps;
  int skip, n_sub;
  int smem_max;       // cudaDevAttrMaxSharedMemoryPerBlockOptin - 1 KB
  size_t smem_bytes;  // dynamic shared memory: the largest role's tables + the slot scratch
  float *slab;
  double *f64;
  unsigned long long *u64;
  int *err;
  cudaStream_t stream;
};

template <int N, int BITS>
cudaError_t role_launch(const RoleLaunchParams &p);

}  // namespace cfb
