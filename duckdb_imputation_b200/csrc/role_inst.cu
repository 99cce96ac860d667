// role_inst.cu -- explicit instantiations of role_scan_kernel for n in [CFB_INST_LO, CFB_INST_HI].
#include "role_kernels.cuh"
#include "role_launch.h"

#ifndef CFB_INST_LO
#error "compile with -DCFB_INST_LO=<n> -DCFB_INST_HI=<n>"
#endif

namespace cfb {

template <int N, int BITS>
cudaError_t role_launch(const RoleLaunchParams &p) {
  RoleArgs a{};
  a.cols = p.cols;
  const Layout &L = *p.lay;
  a.m = L.m;
  a.n_groups = L.n_groups;
  a.U = L.U;
  for (int c = 0; c < kMaxCat; c++) {
    a.lo[c] = L.lo[c];
    a.dom[c] = L.dom[c];
  }
  for (int c = 0; c <= kMaxCat; c++) a.cat_off[c] = (int)L.cat_off[c];
  a.total_dom = L.total_dom;
  a.numcat_base = L.numcat_base;
  a.pair_base = L.pair_base;
  a.plan = *p.plan;
  a.n_rows = p.rows;
  a.chunk_rows = p.chunk_rows;
  a.pair_fold_chunks = p.pair_fold_chunks;
  a.n_reps = p.n_reps;
  a.skip = p.skip;
  a.n_sub = p.n_sub;
  a.slab = p.slab;
  a.f64 = p.f64;
  a.u64 = p.u64;
  a.err = p.err;
  auto kern = role_scan_kernel<N, BITS>;
  // always the device maximum: host threads launch with different sizes concurrently, and a smaller value set by
  // one thread must not undercut another thread's launch
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem_max);
  if (e != cudaSuccess) return e;
  kern<<<p.n_roles * p.n_reps, kRoleThreads, p.smem_bytes, p.stream>>>(a);
  return cudaGetLastError();
}

#define CFB_INST(N)                                                     \
  template cudaError_t role_launch<N, 16>(const RoleLaunchParams &);    \
  template cudaError_t role_launch<N, 32>(const RoleLaunchParams &);
#define CFB_IN_RANGE(N) ((N) >= CFB_INST_LO && (N) <= CFB_INST_HI)

#if CFB_IN_RANGE(0)
CFB_INST(0)
#endif
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
