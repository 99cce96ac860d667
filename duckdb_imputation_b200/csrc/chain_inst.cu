// chain_inst.cu -- explicit instantiations of chain_sum_kernel for n in [CFB_INST_LO, CFB_INST_HI].
#include "chain_kernels.cuh"
#include "chain_launch.h"

#ifndef CFB_INST_LO
#error "compile with -DCFB_INST_LO=<n> -DCFB_INST_HI=<n>"
#endif

namespace cfb {

template <int N>
cudaError_t chain_launch(const ChainLaunchParams &p) {
  ChainArgs a{};
  a.cols = p.cols;
  a.n_rows = p.rows;
  const Layout &L = *p.lay;
  a.m = L.m;
  a.total_dom = (int)L.total_dom;
  a.n_groups = L.n_groups;
  a.F = L.F;
  a.U = L.U;
  a.tile_rows = p.tile_rows;
  a.fold_tiles = p.fold_tiles;
  a.sub_shift = p.sub_shift;
  a.head_cap = p.head_cap;
  a.adaptive = p.adaptive;
  for (int c = 0; c < kMaxCat; c++) {
    a.lo[c] = L.lo[c];
    a.dom[c] = L.dom[c];
  }
  for (int c = 0; c <= kMaxCat; c++) a.cat_off[c] = (int)L.cat_off[c];
  a.numcat_base = L.numcat_base;
  a.slab = p.slab;
  a.cnt_slab = p.cnt_slab;
  a.f64 = p.f64;
  a.u64 = p.u64;
  a.err = p.err;
  a.packed = p.packed;
  a.packed_stride = p.packed_stride;
  auto kern = chain_sum_kernel<N>;
  // always the device maximum: host threads launch with different sizes concurrently, and a smaller value set by
  // one thread must not undercut another thread's launch
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem_max);
  if (e != cudaSuccess) return e;
  kern<<<p.grid, kChainThreads, p.smem_bytes, p.stream>>>(a);
  return cudaGetLastError();
}

#define CFB_INST(N) template cudaError_t chain_launch<N>(const ChainLaunchParams &);
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
