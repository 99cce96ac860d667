// slot_gram_kernel.cuh -- GROUP BY / filtered numeric aggregation, v2: the rows of a tile are SORTED by slot
// in shared memory, then every slot's segment is reduced with 128-bit operand reads into registers.
//
// Same job as group_kernel.cuh (N, lin_agg, quad_agg of every GROUP BY slot in one pass: the per-row
// states[sdata.sel->get_index(j)] routing of sum_no_lift.cpp:83-147 / sum_to_nb_agg.cpp:61-117), for the
// common case of few slots (QDA per-class triples, the MICE observed / NULL split).  group_scan_kernel keeps
// a warp-private fp32 table per slot in shared memory and pays, per row and 32 output entries, two operand
// reads, a table read, an FFMA and a table write.  Here:
//   1. the slots of a tile are ranked WITHOUT atomics: per 32 rows one ballot per slot gives the warp's
//      count and every lane's rank; a block-wide exclusive scan over (slot, step, warp) turns the counts into
//      positions; every slot's segment starts at a multiple of 4 rows and is padded with zero rows;
//   2. the tile is written to shared memory COLUMN-major in sorted order (plus a column of ones: lin_i is
//      the product x_i * 1), so 4 consecutive rows of a column and slot are one 16-byte word;
//   3. warp task (slot, part, split) walks its share of the slot's 4-row groups; a lane owns E 2x2 BLOCKS of the
//      upper triangle over the columns [x_0..x_{n-1}, 1] -- (i0,i1) x (j0,j1) -- and does, per block and 4 rows,
//      four LDS.128 and sixteen FFMA (half the operand reads per FFMA of one-entry-per-lane; the shared-memory
//      pipe is what bounds this kernel).  The task -> warp assignment is static (up to TPW tasks per warp), so
//      the accumulators persist across tiles and are folded into the fp64 state every `fold_tiles` tiles only
//      (bounds every fp32 run); N is the exact segment size.
// Rows with slot < 0 are filtered out (WHERE / MICE NULL filters); slot >= n_groups is an error.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "slab_kernels.cuh"
#include "state_layout.h"

namespace cfb {

constexpr int kSlotThreads = 256;  // several CTAs per SM: some load and sort their tiles while others add up
constexpr int kSlotWarps = kSlotThreads / 32;
constexpr int kSlotMaxSteps = 4;   // rows per thread and tile
constexpr int kSlotMaxGroups = 32;

struct SlotGramArgs {
  ScanCols cols;
  unsigned long long n_rows;
  int n, kind, n_groups;
  int steps;       // tile rows = steps * kSlotThreads
  int parts;       // warps that share the output entries of one slot
  int splits;      // warps that share the rows of one (slot, part)
  int fold_tiles;  // fold the register accumulators into the state every this many tiles of a CTA
  long long F, U;  // per-slot strides of the f64 / u64 state
  double *f64;
  unsigned long long *u64;
  int *err;
};

// Work items of a slot: 2x2 blocks over the block-columns of [x_0..x_{n-1}, 1, (0)], the whole upper triangle (triple
// ring); or one item per column (NB ring: only sum x and sum x^2 of every column)
__host__ __device__ inline int slot_block_cols(int n) { return (n + 2) / 2; }
__host__ __device__ inline int slot_blocks(int n, int kind) {
  const int nbk = slot_block_cols(n);
  return kind == 0 ? nbk * (nbk + 1) / 2 : n;
}
// column pitch in floats: the tile, the padding of every segment to 4 rows, and 4 more so that pitch % 32 == 4
__host__ __device__ inline int slot_pitch(int n_groups, int steps) {
  const int need = steps * kSlotThreads + 4 * n_groups;
  return (need + 31) / 32 * 32 + 4;
}
__host__ __device__ inline size_t slot_smem_bytes(int n, int n_groups, int steps) {
  return (size_t)(n + 2) * slot_pitch(n_groups, steps) * 4 + (size_t)n_groups * steps * kSlotWarps * 4 + 256;
}

template <int E, int TPW>
__global__ void __launch_bounds__(kSlotThreads, E > 1 ? 2 : 4) slot_gram_kernel(const __grid_constant__ SlotGramArgs a) {
  extern __shared__ __align__(16) float slot_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = a.n, G = a.n_groups, U = a.steps, P = slot_pitch(G, U);
  float *xs = slot_smem;                                            // [n + 2][P], column n = ones, column n + 1 = zeros
  unsigned *wc = reinterpret_cast<unsigned *>(xs + (size_t)(n + 2) * P);  // [G][U][warps] counts -> positions
  __shared__ unsigned warp_tot[32];
  __shared__ unsigned raw_off[kSlotMaxGroups + 1];  // first sorted position of the slot before padding
  __shared__ unsigned seg_start[kSlotMaxGroups], seg_rows[kSlotMaxGroups], seg_size[kSlotMaxGroups];
  const int ES = G * U * kSlotWarps;  // scan entries, slot-major

  // this warp's tasks (task t = warp + k * warps, k < TPW; task -> (slot, part, split), split fastest); the
  // block columns of this lane's blocks, packed bi | bj << 8 (-1 = no block)
  const int per_slot = a.parts * a.splits, n_tasks = G * per_slot;
  const int nbk = slot_block_cols(n), NB = slot_blocks(n, a.kind);
  int ops[TPW][E];
  float acc[TPW][E][4][2];  // [di * 2 + dj][half of the 4-row group]
#pragma unroll
  for (int k = 0; k < TPW; k++) {
    const int t = warp + k * kSlotWarps, tpart = (t % per_slot) / a.splits;
#pragma unroll
    for (int e = 0; e < E; e++) {
#pragma unroll
      for (int v = 0; v < 4; v++) acc[k][e][v][0] = acc[k][e][v][1] = 0.f;
      const int b = (tpart * E + e) * 32 + lane;
      ops[k][e] = -1;
      if (t >= n_tasks || b >= NB) continue;
      if (a.kind == 0) {
        int p = b, bi = 0;
        while (p >= nbk - bi) {
          p -= nbk - bi;
          bi++;
        }
        ops[k][e] = bi | ((bi + p) << 8);
      } else {
        ops[k][e] = b;  // NB ring: the column
      }
    }
  }
  // zero column n + 1 once: it pads an odd number of columns to whole blocks and is never written again
  for (int i = tid; i < P; i += kSlotThreads) xs[(size_t)(n + 1) * P + i] = 0.f;
  auto fold = [&]() {
#pragma unroll
    for (int k = 0; k < TPW; k++) {
      const int t = warp + k * kSlotWarps, tg = t / per_slot;
#pragma unroll
      for (int e = 0; e < E; e++) {
        if (ops[k][e] < 0) continue;
        double *f = a.f64 + tg * a.F;  // the f64 state starts with [lin n | quad nq]
        if (a.kind != 0) {  // NB ring: [0] = sum x^2, [1] = sum x of column ops
          const float q2 = acc[k][e][0][0] + acc[k][e][0][1], l1 = acc[k][e][1][0] + acc[k][e][1][1];
          if (q2 != 0.f) atomicAdd(f + n + ops[k][e], (double)q2);
          if (l1 != 0.f) atomicAdd(f + ops[k][e], (double)l1);
          acc[k][e][0][0] = acc[k][e][0][1] = acc[k][e][1][0] = acc[k][e][1][1] = 0.f;
          continue;
        }
        const int bi = ops[k][e] & 255, bj = ops[k][e] >> 8;
#pragma unroll
        for (int v = 0; v < 4; v++) {
          const float sum = acc[k][e][v][0] + acc[k][e][v][1];
          acc[k][e][v][0] = acc[k][e][v][1] = 0.f;
          const int i = 2 * bi + (v >> 1), j = 2 * bj + (v & 1);
          if (sum == 0.f || i > j || j > n || i >= n) continue;  // below the diagonal, padding, or the count (1 x 1)
          if (j == n) atomicAdd(f + i, (double)sum);  // x_i * 1
          else atomicAdd(f + n + ((long long)i * n - (long long)i * (i + 1) / 2 + j), (double)sum);
        }
      }
    }
  };

  const int T = U * kSlotThreads;
  const unsigned long long n_tiles = (a.n_rows + T - 1) / T;
  int since_fold = 0;
  for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const unsigned long long lo = tile * T;
    const int cnt = (int)min((unsigned long long)T, a.n_rows - lo);
    // ---- 1a. slots, per-warp counts and per-lane ranks by ballots
    int slot[kSlotMaxSteps];
    unsigned pos[kSlotMaxSteps];
#pragma unroll
    for (int u = 0; u < kSlotMaxSteps; u++) {
      slot[u] = -1;
      pos[u] = 0;
      const int row = u * kSlotThreads + tid;
      if (u < U && row < cnt) {
        int g = a.cols.group[lo + row];
        if (g >= G) {
          atomicExch(a.err, 2);
          g = -1;
        }
        slot[u] = g;
      }
      if (u >= U) continue;
      for (int g = 0; g < G; g++) {
        const unsigned m = __ballot_sync(0xffffffffu, slot[u] == g);
        if (slot[u] == g) pos[u] = __popc(m & ((1u << lane) - 1));
        if (lane == 0) wc[(g * U + u) * kSlotWarps + warp] = __popc(m);
      }
    }
    __syncthreads();
    // ---- 1b. exclusive scan of the counts (slot-major), 4 consecutive entries per thread
    {
      unsigned x[4], sum = 0;
#pragma unroll
      for (int e = 0; e < 4; e++) {
        x[e] = 4 * tid + e < ES ? wc[4 * tid + e] : 0u;
        sum += x[e];
      }
      unsigned incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      if (lane == 31) warp_tot[warp] = incl;
      __syncthreads();
      if (warp == 0) {
        const unsigned w = lane < kSlotWarps ? warp_tot[lane] : 0u;
        unsigned wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned y = __shfl_up_sync(0xffffffffu, wi, o);
          if (lane >= o) wi += y;
        }
        warp_tot[lane] = wi - w;
        if (lane == 31) raw_off[G] = wi;  // rows of the tile that passed the filter
      }
      __syncthreads();
      unsigned run = warp_tot[warp] + incl - sum;
#pragma unroll
      for (int e = 0; e < 4; e++)
        if (4 * tid + e < ES) {
          const int idx = 4 * tid + e;
          wc[idx] = run;
          if (idx % (U * kSlotWarps) == 0) raw_off[idx / (U * kSlotWarps)] = run;
          run += x[e];
        }
    }
    __syncthreads();
    // ---- 1c. segments: every slot starts at a multiple of 4 rows
    if (tid == 0) {
      unsigned start = 0;
      for (int g = 0; g < G; g++) {
        const unsigned size = raw_off[g + 1] - raw_off[g];
        seg_start[g] = start;
        seg_size[g] = size;
        seg_rows[g] = (size + 3) & ~3u;
        start += seg_rows[g];
      }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kSlotMaxSteps; u++)
      if (slot[u] >= 0) pos[u] += seg_start[slot[u]] + wc[(slot[u] * U + u) * kSlotWarps + warp] - raw_off[slot[u]];
    // ---- 2. the tile, column-major in sorted order; zero rows pad every segment; N is the segment size
    if (tid < G) {
      if (seg_size[tid]) red_u64(a.u64 + tid * a.U, seg_size[tid]);
      for (unsigned r = seg_start[tid] + seg_size[tid]; r < seg_start[tid] + seg_rows[tid]; r++)
        for (int i = 0; i <= n; i++) xs[(size_t)i * P + r] = 0.f;
    }
#pragma unroll 4
    for (int i = 0; i < n; i++) {
      float v[kSlotMaxSteps];
#pragma unroll
      for (int u = 0; u < kSlotMaxSteps; u++) v[u] = slot[u] >= 0 ? a.cols.num[i][lo + u * kSlotThreads + tid] : 0.f;
#pragma unroll
      for (int u = 0; u < kSlotMaxSteps; u++)
        if (slot[u] >= 0) xs[(size_t)i * P + pos[u]] = v[u];
    }
#pragma unroll
    for (int u = 0; u < kSlotMaxSteps; u++)
      if (slot[u] >= 0) xs[(size_t)n * P + pos[u]] = 1.f;
    __syncthreads();
    // ---- 3. this warp's shares of its slots' 4-row groups
#pragma unroll
    for (int k = 0; k < TPW; k++) {
      const int t = warp + k * kSlotWarps;
      if (t >= n_tasks) continue;
      const int tg = t / per_slot, tsplit = t % a.splits;
      const unsigned groups = seg_rows[tg] / 4;
      const float4 *base = reinterpret_cast<const float4 *>(xs + seg_start[tg]);
      const int P4 = P / 4;
      const float4 *xi[E], *xj[E];  // first column of the block's column pairs, at the first row of the segment
#pragma unroll
      for (int e = 0; e < E; e++) {
        const int o = ops[k][e] < 0 ? 0 : ops[k][e];  // no block: block (0,0), never folded
        xi[e] = base + (size_t)(a.kind == 0 ? 2 * (o & 255) : o) * P4;
        xj[e] = base + (size_t)(2 * (o >> 8)) * P4;
      }
      if (a.kind != 0) {  // NB ring: one load per column and 4 rows
#pragma unroll 2
        for (unsigned q = tsplit; q < groups; q += a.splits) {
#pragma unroll
          for (int e = 0; e < E; e++) {
            const float4 x = xi[e][q];
            acc[k][e][0][0] = fmaf(x.x, x.x, fmaf(x.y, x.y, acc[k][e][0][0]));
            acc[k][e][0][1] = fmaf(x.z, x.z, fmaf(x.w, x.w, acc[k][e][0][1]));
            acc[k][e][1][0] += x.x + x.y;
            acc[k][e][1][1] += x.z + x.w;
          }
        }
        continue;
      }
#pragma unroll 2
      for (unsigned q = tsplit; q < groups; q += a.splits) {
#pragma unroll
        for (int e = 0; e < E; e++) {
          const float4 i0 = xi[e][q], i1 = xi[e][P4 + q], j0 = xj[e][q], j1 = xj[e][P4 + q];
          acc[k][e][0][0] = fmaf(i0.x, j0.x, fmaf(i0.y, j0.y, acc[k][e][0][0]));
          acc[k][e][0][1] = fmaf(i0.z, j0.z, fmaf(i0.w, j0.w, acc[k][e][0][1]));
          acc[k][e][1][0] = fmaf(i0.x, j1.x, fmaf(i0.y, j1.y, acc[k][e][1][0]));
          acc[k][e][1][1] = fmaf(i0.z, j1.z, fmaf(i0.w, j1.w, acc[k][e][1][1]));
          acc[k][e][2][0] = fmaf(i1.x, j0.x, fmaf(i1.y, j0.y, acc[k][e][2][0]));
          acc[k][e][2][1] = fmaf(i1.z, j0.z, fmaf(i1.w, j0.w, acc[k][e][2][1]));
          acc[k][e][3][0] = fmaf(i1.x, j1.x, fmaf(i1.y, j1.y, acc[k][e][3][0]));
          acc[k][e][3][1] = fmaf(i1.z, j1.z, fmaf(i1.w, j1.w, acc[k][e][3][1]));
        }
      }
    }
    if (++since_fold >= a.fold_tiles || tile + gridDim.x >= n_tiles) {
      since_fold = 0;
      fold();
    }
    __syncthreads();  // the tile buffers are rewritten by the next iteration
  }
}

}  // namespace cfb
