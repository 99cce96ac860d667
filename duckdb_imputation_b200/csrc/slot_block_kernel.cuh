// slot_block_kernel.cuh -- GROUP BY / filtered numeric aggregation, third generation: the tile is fetched by bulk
// async copies, sorted by slot in shared memory, and reduced with 4x4 REGISTER BLOCKS and packed FFMA2.
//
// Same job as slot_gram_kernel.cuh (N, lin_agg, quad_agg of every GROUP BY slot in one pass: the per-row
// states[sdata.sel->get_index(j)] routing of sum_no_lift.cpp:83-147 / sum_to_nb_agg.cpp:61-117), for <= 32 slots.
// What ncu showed on slot_gram_kernel and on the first version of this file, and what changed:
//   * fetch: ordinary global loads left the sorted stores waiting on long_scoreboard and the other warps at the
//     barriers behind them (45 % of the samples).  Now one warp issues 1-D bulk async copies (cp.async.bulk, TMA
//     engine, completion on an mbarrier) of the NEXT tile's raw columns while the current tile is multiplied; the sort
//     reads shared memory.
//   * sort: ranking rows with one ballot per (row step, slot) costs 80 instructions per row; here a row takes its rank
//     from ONE shared-memory atomicAdd on its slot's counter, and one warp turns the counters into 4-row-aligned
//     segments (no block-wide scan).
//   * multiply: a lane owns a 4x4 block of the upper triangle over [x_0..x_{n-1}] -- eight 128-bit operand reads per
//     64 FMAs, issued as 32 fma.rn.f32x2 (two rows each): a quarter of the operand traffic per FMA of the 2x2 blocks.
//     The column sums (lin_agg) ride along as packed adds instead of a column of ones (n = 12 is 3 block columns, not
//     4).  The lanes of a warp that share a block column read one address (a broadcast wavefront); the columns are
//     skewed so that block columns fall into different bank groups.
// The work of a slot is cut into (block, split) lane tasks, statically assigned: the accumulators live in registers
// across tiles and are folded into the fp64 state every `fold_tiles` tiles (bounds every fp32 run).
// NB = true is the Naive-Bayes ring (sum x and sum x^2 per column): the "block" is a group of 4 columns.
// Rows with slot < 0 are filtered out (WHERE / MICE NULL filters); slot >= n_groups is an error.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "ptx_sm100.cuh"
#include "slab_kernels.cuh"
#include "slot_gram_kernel.cuh"
#include "state_layout.h"

namespace cfb {

struct SlotBlockArgs {
  ScanCols cols;
  unsigned long long n_rows;
  int n, n_groups;
  int steps;       // tile rows = steps * kSlotThreads
  int splits;      // lanes that share the rows of one (slot, block)
  int fold_tiles;  // fold the register accumulators into the state every this many tiles of a CTA
  long long F, U;  // per-slot strides of the f64 / u64 state
  double *f64;
  unsigned long long *u64;
  int *err;
  // Naive-Bayes ring only: the key counts of the m categorical columns (sum_to_nb_agg.cpp:124-145) in the same pass --
  // their columns are fetched with the tile and counted in a shared-memory histogram over (slot, column, key),
  // folded into the u64 state when the CTA is done (m == 0: not fused, key_count_kernel does them)
  int m, total_dom;
  int lo[kMaxCat], dom[kMaxCat], cat_off[kMaxCat + 1];
};

// columns of the sorted tile: [x_0..x_{n-1}, zero padding] in whole blocks of 4
__host__ __device__ inline int slotb_block_cols(int n) { return (n + 3) / 4; }
__host__ __device__ inline int slotb_blocks(int n, bool nb_ring) {
  const int nb = slotb_block_cols(n);
  return nb_ring ? nb : nb * (nb + 1) / 2;
}
// quads (4 rows) per column: the tile, the padding of every segment to 4 rows, rounded to a multiple of 8 quads plus
// 8 quads of room for the skew
__host__ __device__ inline int slotb_pitch_quads(int n_groups, int steps) {
  const int need = (steps * kSlotThreads + 4 * n_groups + 3) / 4;
  return (need + 7) / 8 * 8 + 8;
}
// first quad of column c: columns r, 4 + r, 8 + r, 12 + r (what the lanes of a warp read together) get the bank
// groups r, r + 2, r + 4, r + 6
__host__ __device__ inline int slotb_col_quad(int c, int pitch_quads) { return c * pitch_quads + (((c & 3) + 2 * (c >> 2)) & 7); }
// dynamic shared memory: sorted tile | raw staging [n + 1 + m][tile rows] (column n = the slot ids, then the m key
// columns of a fused Naive-Bayes scan) | key histogram [n_groups * total_dom]
__host__ __device__ inline size_t slotb_smem_bytes(int n, int n_groups, int steps, int m = 0, int total_dom = 0) {
  return (size_t)slotb_block_cols(n) * 4 * slotb_pitch_quads(n_groups, steps) * 16 +
         (size_t)(n + 1 + m) * steps * kSlotThreads * 4 + (size_t)(m ? n_groups * total_dom : 0) * 4;
}

struct __align__(16) SlotQuad {
  unsigned long long lo, hi;  // rows (r, r+1) and (r+2, r+3) of one column, as packed float pairs
};
__device__ __forceinline__ void slotb_ffma2(unsigned long long &d, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ void slotb_fadd2(unsigned long long &d, unsigned long long a) {
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(d) : "l"(a));
}
__device__ __forceinline__ float slotb_pair_sum(unsigned long long v) {
  const float2 p = *reinterpret_cast<const float2 *>(&v);
  return p.x + p.y;
}

template <bool NB>
__global__ void __launch_bounds__(kSlotThreads, 2) slot_block_kernel(const __grid_constant__ SlotBlockArgs a) {
  extern __shared__ __align__(16) float slotb_smem[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int n = a.n, G = a.n_groups, U = a.steps, PQ = slotb_pitch_quads(G, U), T = U * kSlotThreads;
  const int nb = slotb_block_cols(n), NC = 4 * nb, NBLK = NB ? nb : nb * (nb + 1) / 2;
  float *xs = slotb_smem;                       // [NC] columns of PQ quads, skewed; sorted by slot
  float *raw = xs + (size_t)NC * PQ * 4;        // [n + 1 + m][T]: the tile as it lies in the table, column n = slot ids
  const int m = NB ? a.m : 0, n_fetch = n + 1 + m;
  unsigned *key_hist = reinterpret_cast<unsigned *>(raw + (size_t)n_fetch * T);  // [G * total_dom] (fused NB scan)
  for (int i = tid; i < (m ? G * a.total_dom : 0); i += kSlotThreads) key_hist[i] = 0;
  __shared__ __align__(8) uint64_t full_bar;
  __shared__ unsigned cnt[2][kSlotMaxGroups];   // rows of every slot in the tile; tiles alternate between the two sets

  // this lane's task: (slot, split, block), block fastest -- the lanes of a warp mostly share slot and split, i.e. rows
  const int per_slot = NBLK * a.splits;
  const bool active = tid < G * per_slot;
  const int tg = active ? tid / per_slot : 0, tsplit = active ? (tid % per_slot) / NBLK : 0;
  int bi = 0, bj = 0;
  {
    int p = active ? tid % NBLK : 0;
    if (NB) {
      bi = bj = p;
    } else {
      while (p >= nb - bi) {
        p -= nb - bi;
        bi++;
      }
      bj = bi + p;
    }
  }
  int qi[4], qj[4];  // first quad of the block's columns
#pragma unroll
  for (int r = 0; r < 4; r++) {
    qi[r] = slotb_col_quad(4 * bi + r, PQ);
    qj[r] = slotb_col_quad(4 * bj + r, PQ);
  }
  // triple ring: acc[r * 4 + c] = packed partial sums of x_{4bi+r} * x_{4bj+c} over even / odd row pairs;
  // NB ring: acc[r] = sums of x_{4bi+r}^2.  lin[r] = packed partial sums of x_{4bi+r}.
  unsigned long long acc[NB ? 4 : 16], lin[4];
#pragma unroll
  for (int v = 0; v < (NB ? 4 : 16); v++) acc[v] = 0ull;
#pragma unroll
  for (int r = 0; r < 4; r++) lin[r] = 0ull;
  // the padding columns (n .. NC - 1) are zero and never written again
  for (int c = n; c < NC; c++)
    for (int i = tid; i < PQ * 4; i += kSlotThreads) xs[(size_t)c * PQ * 4 + i] = 0.f;
  if (tid < 2 * kSlotMaxGroups) cnt[tid / kSlotMaxGroups][tid % kSlotMaxGroups] = 0;

  auto fold = [&]() {
    if (!active) return;
    double *f = a.f64 + tg * a.F;  // the f64 state starts with [lin n | quad nq]
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int i = 4 * bi + r;
      const float s = slotb_pair_sum(lin[r]);
      lin[r] = 0ull;
      if (bi == bj && i < n && s != 0.f) atomicAdd(f + i, (double)s);
    }
#pragma unroll
    for (int v = 0; v < (NB ? 4 : 16); v++) {
      const float sum = slotb_pair_sum(acc[v]);
      acc[v] = 0ull;
      const int i = 4 * bi + (NB ? v : (v >> 2)), j = NB ? i : 4 * bj + (v & 3);
      if (sum == 0.f || i > j || j >= n) continue;  // below the diagonal or padding
      if (NB) atomicAdd(f + n + i, (double)sum);
      else atomicAdd(f + n + ((long long)i * n - (long long)i * (i + 1) / 2 + j), (double)sum);
    }
  };

  const unsigned long long n_tiles = (a.n_rows + T - 1) / T;
  // Fetch of one tile (warp 0): column c < n and the slot column (c == n), whole 16-byte pieces by bulk async copies,
  // the <= 3 rows of a ragged tail by plain loads (only the table's last tile has one).
  auto fetch = [&](unsigned long long tile) {
    const unsigned long long lo = tile * T;
    const int cnt_rows = (int)min((unsigned long long)T, a.n_rows - lo), whole = cnt_rows & ~3;
    if (tid == 0) ptx::mbar_arrive_expect_tx(&full_bar, (uint32_t)(whole * 4 * n_fetch));
    __syncwarp();
    if (tid < n_fetch) {
      const void *src = tid < n ? (const void *)(a.cols.num[tid] + lo)
                                : (tid == n ? (const void *)(a.cols.group + lo) : (const void *)(a.cols.cat[tid - n - 1] + lo));
      if (whole) ptx::bulk_g2s_plain(raw + (size_t)tid * T, src, (uint32_t)(whole * 4), &full_bar);
      for (int r = whole; r < cnt_rows; r++) raw[(size_t)tid * T + r] = reinterpret_cast<const float *>(src)[r];
    }
  };
  if (tid == 0) {
    ptx::mbar_init(&full_bar, 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  if (blockIdx.x < n_tiles && tid < 32) fetch(blockIdx.x);
  int since_fold = 0;
  unsigned it = 0;
  for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
    const unsigned long long lo = tile * T;
    const int rows_here = (int)min((unsigned long long)T, a.n_rows - lo);
    ptx::mbar_wait(&full_bar, it & 1);
    if (rows_here & 3) __syncthreads();  // ragged tail: the plain stores of warp 0 must be visible to everyone
    const int *rslot = reinterpret_cast<const int *>(raw + (size_t)n * T);
    // ---- 1. every row takes its rank within its slot from the slot's counter
    unsigned *tile_cnt = cnt[it & 1];
    int slot[kSlotMaxSteps];
    unsigned pos[kSlotMaxSteps];
#pragma unroll
    for (int u = 0; u < kSlotMaxSteps; u++) {
      slot[u] = -1;
      pos[u] = 0;
      const int row = u * kSlotThreads + tid;
      if (u < U && row < rows_here) {
        int g = rslot[row];
        if (g >= G) {
          atomicExch(a.err, 2);
          g = -1;
        }
        slot[u] = g;
        if (g >= 0) pos[u] = atomicAdd(&tile_cnt[g], 1u);
        if (NB && g >= 0) {
          for (int c = 0; c < m; c++) {
            const unsigned sl = (unsigned)(reinterpret_cast<const int *>(raw + (size_t)(n + 1 + c) * T)[row] - a.lo[c]);
            if (sl < (unsigned)a.dom[c]) atomicAdd(&key_hist[g * a.total_dom + a.cat_off[c] + (int)sl], 1u);
            else atomicExch(a.err, 1);  // a key outside the declared domain
          }
        }
      }
    }
    __syncthreads();
    // ---- 2. segments: every slot starts at a multiple of 4 rows and is padded with zero rows.  Every WARP derives
    //         them for itself (lane g holds slot g's start and size): no shared arrays, no barrier of their own.
    const unsigned seg_size = lane < G ? tile_cnt[lane] : 0u, seg_rows4 = (seg_size + 3) & ~3u;
    unsigned seg_begin;
    {
      unsigned incl = seg_rows4;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      seg_begin = incl - seg_rows4;
    }
    if (tid < G && seg_size) red_u64(a.u64 + tid * a.U, seg_size);  // N += size
    cnt[(it & 1) ^ 1][lane] = 0;  // the other counter set (last read before the barrier above) serves the next tile
    // ---- 3. the tile, column-major in sorted order; the zero rows that pad a segment to whole 4-row groups
#pragma unroll
    for (int u = 0; u < kSlotMaxSteps; u++) {
      const unsigned st = __shfl_sync(0xffffffffu, seg_begin, slot[u] < 0 ? 0 : slot[u]);
      if (slot[u] >= 0) pos[u] += st;
    }
    for (int g = tid >> 5; g < G; g += kSlotWarps) {  // warp w pads the slots w, w + 8, ...: lane = (pad row, column)
      const unsigned st = __shfl_sync(0xffffffffu, seg_begin, g), sz = __shfl_sync(0xffffffffu, seg_size, g);
      const unsigned pad = ((sz + 3) & ~3u) - sz;
      for (unsigned e = lane; e < pad * (unsigned)n; e += 32)
        xs[(size_t)slotb_col_quad((int)(e % n), PQ) * 4 + st + sz + e / n] = 0.f;
    }
#pragma unroll 4
    for (int i = 0; i < n; i++) {
      float v[kSlotMaxSteps];
#pragma unroll
      for (int u = 0; u < kSlotMaxSteps; u++) v[u] = slot[u] >= 0 ? raw[(size_t)i * T + u * kSlotThreads + tid] : 0.f;
      float *col = xs + (size_t)slotb_col_quad(i, PQ) * 4;
#pragma unroll
      for (int u = 0; u < kSlotMaxSteps; u++)
        if (slot[u] >= 0) col[pos[u]] = v[u];
    }
    __syncthreads();
    // the staging buffer has been read: the next tile of this CTA lands in it while this one is multiplied
    if (tile + gridDim.x < n_tiles && tid < 32) fetch(tile + gridDim.x);
    // ---- 4. this lane's block over its share of the slot's 4-row groups
    const unsigned my_rows4 = __shfl_sync(0xffffffffu, seg_rows4, tg), my_begin = __shfl_sync(0xffffffffu, seg_begin, tg);
    if (active) {
      const unsigned groups = my_rows4 / 4;
      const SlotQuad *base = reinterpret_cast<const SlotQuad *>(xs) + my_begin / 4;
#pragma unroll 1
      for (unsigned q = tsplit; q < groups; q += a.splits) {
        SlotQuad xi[4];
#pragma unroll
        for (int r = 0; r < 4; r++) {
          xi[r] = base[qi[r] + q];
          slotb_fadd2(lin[r], xi[r].lo);
          slotb_fadd2(lin[r], xi[r].hi);
        }
        if constexpr (NB) {
#pragma unroll
          for (int r = 0; r < 4; r++) {
            slotb_ffma2(acc[r], xi[r].lo, xi[r].lo);
            slotb_ffma2(acc[r], xi[r].hi, xi[r].hi);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 4; c++) {
            const SlotQuad xj = base[qj[c] + q];
#pragma unroll
            for (int r = 0; r < 4; r++) {
              slotb_ffma2(acc[r * 4 + c], xi[r].lo, xj.lo);
              slotb_ffma2(acc[r * 4 + c], xi[r].hi, xj.hi);
            }
          }
        }
      }
    }
    if (++since_fold >= a.fold_tiles || tile + gridDim.x >= n_tiles) {
      since_fold = 0;
      fold();
    }
    __syncthreads();  // the sorted tile is rewritten by the next iteration
  }
  // key counts of the fused Naive-Bayes scan (a CTA sees < 2^32 rows)
  for (int i = tid; i < (m ? G * a.total_dom : 0); i += kSlotThreads) {
    const unsigned v = key_hist[i];
    if (v) red_u64(a.u64 + (i / a.total_dom) * a.U + 1 + i % a.total_dom, v);
  }
}

}  // namespace cfb
