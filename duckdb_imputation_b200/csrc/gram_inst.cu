// gram_inst.cu -- explicit instantiations of the Gram kernels for n in [CFB_INST_LO, CFB_INST_HI].
#include <algorithm>
#include <mutex>

#include "gram_kernel.cuh"
#include "gram_launch.h"

#ifndef CFB_INST_LO
#error "compile with -DCFB_INST_LO=<n> -DCFB_INST_HI=<n>"
#endif

namespace cfb {

template <int N, bool DIAG>
cudaError_t gram_launch(const GramLaunchParams &p) {
  using S = GramShape<N, DIAG>;
  constexpr int TR = GramTile<N, DIAG>::kRows;
  const size_t fixed = gram_smem_bytes<N, DIAG>(0);
  const size_t per_stage = (size_t)N * TR * sizeof(float) + 2 * sizeof(uint64_t);
  if ((size_t)p.smem_optin < fixed + 2 * per_stage + 1024) return cudaErrorInvalidValue;
  int stages = (int)std::min<size_t>(8, ((size_t)p.smem_optin - 1024 - fixed) / per_stage);
  if (p.stages > 0) stages = std::max(2, std::min(stages, p.stages));
  const size_t smem = gram_smem_bytes<N, DIAG>(stages);
  auto kern = gram_scan_kernel<N, DIAG, TR>;
  static std::once_flag once[64];
  cudaError_t attr_err = cudaSuccess;
  std::call_once(once[p.device & 63], [&] {
    attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem_optin - 1024);
  });
  if (attr_err != cudaSuccess) return attr_err;
  GramArgs a{};
  for (int k = 0; k < N; k++) a.cols.p[k] = p.cols[k];
  a.n_rows = p.rows;
  a.stages = stages;
  // default: an fp32 accumulator half sees at most ~256 rows between folds into fp64
  a.flush_tiles = p.flush_tiles > 0 ? p.flush_tiles : std::max(1, 256 * 64 * S::kGroups / TR);
  a.partials = p.partials;
  a.state = p.state;
  a.ticket = p.ticket;
  a.count = p.count;
  const unsigned long long tiles = ((p.rows & ~3ull) + TR - 1) / TR;
  const int grid = (int)std::max<unsigned long long>(1, std::min<unsigned long long>(p.max_grid, tiles));
  kern<<<grid, S::kThreads, smem, p.stream>>>(a);
  return cudaGetLastError();
}

#define CFB_INST(N)                                                        \
  template cudaError_t gram_launch<N, false>(const GramLaunchParams &);    \
  template cudaError_t gram_launch<N, true>(const GramLaunchParams &);

#define CFB_IN_RANGE(N) ((N) >= CFB_INST_LO && (N) <= CFB_INST_HI)

#if CFB_IN_RANGE(1)
CFB_INST(1)
#endif
#if CFB_IN_RANGE(2)
CFB_INST(2)
#endif
#if CFB_IN_RANGE(3)
CFB_INST(3)
#endif
#if CFB_IN_RANGE(4)
CFB_INST(4)
#endif
#if CFB_IN_RANGE(5)
CFB_INST(5)
#endif
#if CFB_IN_RANGE(6)
CFB_INST(6)
#endif
#if CFB_IN_RANGE(7)
CFB_INST(7)
#endif
#if CFB_IN_RANGE(8)
CFB_INST(8)
#endif
#if CFB_IN_RANGE(9)
CFB_INST(9)
#endif
#if CFB_IN_RANGE(10)
CFB_INST(10)
#endif
#if CFB_IN_RANGE(11)
CFB_INST(11)
#endif
#if CFB_IN_RANGE(12)
CFB_INST(12)
#endif
#if CFB_IN_RANGE(13)
CFB_INST(13)
#endif
#if CFB_IN_RANGE(14)
CFB_INST(14)
#endif
#if CFB_IN_RANGE(15)
CFB_INST(15)
#endif
#if CFB_IN_RANGE(16)
CFB_INST(16)
#endif
#if CFB_IN_RANGE(17)
CFB_INST(17)
#endif
#if CFB_IN_RANGE(18)
CFB_INST(18)
#endif
#if CFB_IN_RANGE(19)
CFB_INST(19)
#endif
#if CFB_IN_RANGE(20)
CFB_INST(20)
#endif
#if CFB_IN_RANGE(21)
CFB_INST(21)
#endif
#if CFB_IN_RANGE(22)
CFB_INST(22)
#endif
#if CFB_IN_RANGE(23)
CFB_INST(23)
#endif
#if CFB_IN_RANGE(24)
CFB_INST(24)
#endif
#if CFB_IN_RANGE(25)
CFB_INST(25)
#endif
#if CFB_IN_RANGE(26)
CFB_INST(26)
#endif
#if CFB_IN_RANGE(27)
CFB_INST(27)
#endif
#if CFB_IN_RANGE(28)
CFB_INST(28)
#endif
#if CFB_IN_RANGE(29)
CFB_INST(29)
#endif
#if CFB_IN_RANGE(30)
CFB_INST(30)
#endif
#if CFB_IN_RANGE(31)
CFB_INST(31)
#endif
#if CFB_IN_RANGE(32)
CFB_INST(32)
#endif

}  // namespace cfb
