// gram_launch.h -- host-side launch interface of the Gram kernels (gram_kernel.cuh).
// The 64 kernel instantiations (n = 1..32, triple / NB) are compiled in separate
// translation units (gram_inst.cu with -DCFB_INST_LO/-DCFB_INST_HI) so the build runs in
// parallel; this header is what cofactor_b200.cu sees of them.
#pragma once
#include <cuda_runtime.h>

namespace cfb {

struct GramLaunchParams {
  const float *const *cols;  // n device column pointers, 16-byte aligned
  unsigned long long rows;
  int smem_optin;   // cudaDevAttrMaxSharedMemoryPerBlockOptin of the device
  int max_grid;     // persistent CTAs (normally the SM count)
  int stages;       // 0 = as deep as shared memory allows (<= 8)
  int flush_tiles;  // 0 = default bounded run
  double *partials;
  double *state;
  unsigned int *ticket;
  unsigned long long *count;  // N += rows along with the sums (nullptr: not wanted)
  cudaStream_t stream;
  int device;
};

// Returns cudaSuccess or the launch error; cudaErrorInvalidValue if the ring does not fit.
template <int N, bool DIAG>
cudaError_t gram_launch(const GramLaunchParams &p);

}  // namespace cfb
