// group_kernel.cuh -- GROUP BY / filtered numeric aggregation: N, lin_agg and quad_agg of every
// GROUP BY slot in one pass (the per-row states[sdata.sel->get_index(j)] routing of
// sum_no_lift.cpp:83-147 and sum_to_nb_agg.cpp:61-117).
//
// A row's contribution goes to the accumulators of ITS slot, so register accumulators (the Gram
// kernel) do not apply, and one atomic per value (slab kernel) is L2-bound.  Here the lanes of a
// warp own the OUTPUT entries instead of the rows:
//   * every output of a slot is a product of two columns of the row extended by a constant 1:
//     N = 1*1, lin_i = x_i*1, quad_ij = x_i*x_j  -> V = 1 + n + nq entries, lane L owns entries
//     L, L+32, L+64, ... (E = ceil(V/32) per lane);
//   * each warp keeps a private fp32 table [n_groups][E*32] in shared memory; lane L only ever
//     touches its own entries, so rows are added with plain LDS / FFMA / STS -- no atomics, no bank
//     conflicts (consecutive lanes, consecutive words);
//   * a warp stages 32 rows at a time: coalesced 128-byte loads per column, stored transposed
//     (row-major, odd pitch) so that the per-row operand reads are conflict-free broadcasts;
//   * slot < 0 means "row filtered out" (MICE scans WHERE col_IS_NULL IS FALSE);
//   * tables are folded into the fp64 / u64 state with atomics every `flush_rows` rows per warp,
//     which bounds every fp32 run.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "state_layout.h"

namespace cfb {

constexpr int kGroupWarps = 8;
constexpr int kGroupThreads = kGroupWarps * 32;

struct GroupArgs {
  ScanCols cols;
  unsigned long long n_rows;
  int n, kind, n_groups;
  int flush_rows;       // per warp
  long long F, U;       // per-group strides of the state arrays
  double *f64;
  unsigned long long *u64;
  int *err;
};

__host__ __device__ inline int group_entries(int n, int kind) { return 1 + n + (kind == 0 ? n * (n + 1) / 2 : n); }
// shared memory of one CTA: per warp [n_groups][E*32] floats + [32][pitch] floats + 32 ints
__host__ __device__ inline size_t group_smem_bytes(int n, int n_groups, int E) {
  const int pitch = (n + 1) | 1;
  return (size_t)kGroupWarps * ((size_t)n_groups * E * 32 + 32 * pitch + 32) * 4;
}

template <int E>
__global__ void __launch_bounds__(kGroupThreads) group_scan_kernel(const __grid_constant__ GroupArgs a) {
  extern __shared__ __align__(16) float g_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = a.n, G = a.n_groups;
  const int V = group_entries(n, a.kind);
  const int pitch = (n + 1) | 1;  // odd: transposed stores are conflict-free
  const size_t per_warp = (size_t)G * E * 32 + 32 * pitch + 32;
  float *table = g_smem + warp * per_warp;      // [G][E*32]
  float *tile = table + (size_t)G * E * 32;     // [32][pitch], column n = 1.0
  int *slots = reinterpret_cast<int *>(tile + 32 * pitch);

  // operand columns of this lane's entries (column n is the constant 1)
  int ia[E], ib[E];
#pragma unroll
  for (int e = 0; e < E; e++) {
    const int idx = e * 32 + lane;
    int i = n, j = n;
    if (idx >= 1 && idx <= n) {
      i = idx - 1;
    } else if (idx > n && idx < V) {
      int p = idx - 1 - n;
      if (a.kind == 0) {
        i = 0;
        while (p >= n - i) {
          p -= n - i;
          i++;
        }
        j = i + p;
      } else {
        i = j = p;
      }
    }
    ia[e] = i;
    ib[e] = j;
  }
  for (size_t t = lane; t < (size_t)G * E * 32; t += 32) table[t] = 0.f;
  tile[lane * pitch + n] = 1.0f;
  __syncwarp();

  auto flush = [&]() {
    for (int g = 0; g < G; g++) {
#pragma unroll
      for (int e = 0; e < E; e++) {
        const int idx = e * 32 + lane;
        float *p = table + ((size_t)g * E + e) * 32 + lane;
        const float v = *p;
        if (idx < V && v != 0.f) {
          if (idx == 0)
            atomicAdd(a.u64 + (long long)g * a.U, (unsigned long long)v);
          else
            atomicAdd(a.f64 + (long long)g * a.F + (idx - 1), (double)v);
          *p = 0.f;
        }
      }
    }
    __syncwarp();
  };

  const unsigned long long n_blocks32 = (a.n_rows + 31) / 32;
  const unsigned long long stride = (unsigned long long)gridDim.x * kGroupWarps;
  int since_flush = 0;
  for (unsigned long long blk = (unsigned long long)blockIdx.x * kGroupWarps + warp; blk < n_blocks32; blk += stride) {
    const unsigned long long r = blk * 32 + lane;
    const bool in = r < a.n_rows;
    // stage 32 rows: coalesced column loads, transposed store
    for (int c = 0; c < n; c++) tile[lane * pitch + c] = in ? a.cols.num[c][r] : 0.f;
    int s = in ? (a.cols.group ? a.cols.group[r] : 0) : -1;
    if (s >= G) {
      atomicExch(a.err, 2);
      s = -1;
    }
    slots[lane] = s;
    __syncwarp();
    const int valid = (int)min(32ull, a.n_rows - blk * 32);
    for (int rr = 0; rr < valid; rr++) {
      const int g = slots[rr];
      if (g < 0) continue;  // filtered row
      const float *x = tile + rr * pitch;
      float *acc = table + (size_t)g * E * 32 + lane;
#pragma unroll
      for (int e = 0; e < E; e++) acc[e * 32] = fmaf(x[ia[e]], x[ib[e]], acc[e * 32]);
    }
    __syncwarp();
    since_flush += 32;
    if (since_flush >= a.flush_rows) {
      flush();
      since_flush = 0;
    }
  }
  flush();
}

}  // namespace cfb
