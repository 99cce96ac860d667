// chain_kernels.cuh -- kernel family K3, second generation of the categorical scan (round 2).
//
// Replaces the per-row std::map updates of Triple::SumNoLift (sum_no_lift.cpp:158-214):
//   chain_sum_kernel<N>    key counts + per-key sums [count, sum x_0..x_{n-1}] of every categorical column
//                          (lin_cat / quad_num_cat), and -- fused into the same pass over the keys -- the
//                          validated, packed one-byte slots of every row for the pair kernel;
//   pair_packed_kernel     (key1,key2) pair counts (quad_cat) in shared-memory tables from those packed slots.
// Together they read every input byte from HBM once (the int32 keys are never loaded twice) and write / re-read
// (m+1) bytes per row of packed slots.
//
// chain_sum_kernel, per tile of T rows in shared memory (no float atomics anywhere):
//   1. one pass over the tile's rows: the numeric values of a row are stored as its payload row
//      [x_0..x_{n-1}] (16-byte quads, odd quad stride: conflict-free stores); every (row, column) is pushed on
//      the linked list of its bucket (bucket = (slot, column, key) x S sub-lists by row) with ONE integer
//      ATOMS.EXCH on the list head; the previous head becomes the row's `next` (a coalesced 16-bit store).
//      The first generation (bucket_kernels.cuh) counted, scanned and then ranked every (row, column) again
//      with a second ATOMS and a second load of the key; the lists need neither.
//   2. list walk: Q adjacent lanes per list (one per 16-byte quad of the payload; the lanes of a team read one
//      payload row as contiguous pieces) follow head -> next -> ... and add the payload rows in registers; the
//      list length is the key count (the payload carries no count column: n = 20 takes 5 quads, not 6).  The
//      result goes to the CTA's fp32 slab in L2 as one 128-bit vector reduction per (list, quad) -- fire and
//      forget, a few per row instead of 3 per (row, column);
//   3. every `fold_tiles` tiles (<= ~32 K rows: bounds every fp32 run) the slab is folded into the fp64 / u64 state.
// Skewed keys: a hot key's list would keep one team of lanes busy long after every other list has ended, so the
// number of sub-lists of a bucket follows the bucket's length in the CTA's previous tile (~32 rows per sub-list, heads
// handed out by a prefix sum; a head's upper half-word names its bucket); as long as no bucket needs more than twice the
// static plan the static plan (uniform shift, no lookups) stays in force.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "slab_kernels.cuh"
#include "role_kernels.cuh"
#include "state_layout.h"

namespace cfb {

constexpr int kChainThreads = 1024;
constexpr int kChainCtasPerSm = 1;
constexpr int kChainMaxHeads = 4096;    // buckets x sub-lists: one u32 head each in shared memory
constexpr unsigned kChainEnd = 0xFFFFu;  // end of list (row ids are 16-bit: tile_rows <= 65535)

struct ChainArgs {
  ScanCols cols;
  unsigned long long n_rows;
  int m, total_dom;
  int n_groups;    // GROUP BY slots: bucket = (slot, column, key)
  long long F, U;  // per-slot strides of the f64 / u64 state
  int tile_rows;   // multiple of 32, <= 65504
  int fold_tiles;  // fold the slab into the state every this many tiles of a CTA
  int sub_shift;   // S = 1 << sub_shift sub-lists per bucket (keeps all lanes busy when there are few buckets)
  int head_cap;    // list heads in shared memory (>= D << sub_shift, <= kChainMaxHeads)
  int adaptive;    // 0: always the static plan (CFB_CHAIN_NO_ADAPT: measurements)
  int lo[kMaxCat], dom[kMaxCat], cat_off[kMaxCat + 1];
  long long numcat_base;
  float *slab;         // [gridDim.x][D * 4Q] fp32 sums, all zero on entry and on exit
  unsigned *cnt_slab;  // [gridDim.x][D] counts, all zero on entry and on exit
  double *f64;
  unsigned long long *u64;
  int *err;
  // packed slots for pair_packed_kernel (nullptr: not wanted): column c of row r at packed[c * stride + r].  A scan
  // with a GROUP BY / filter column has two more columns: m = the row's slot (0 when filtered), m + 1 = 1 for a live
  // row, 0 for a filtered one (its increment), and its rows in [n_rows, round-up-to-16) are written as filtered.  A key
  // outside the declared domain is packed as slot 0 (the scan fails with CFB_ERR_DOMAIN anyway).
  unsigned char *packed;
  unsigned long long packed_stride;
};

__host__ __device__ constexpr int chain_quads(int n) { return (n + 3) / 4; }
__host__ __device__ constexpr int chain_quad_stride(int n) { return chain_quads(n) | 1; }  // odd: conflict-free payload stores

// dynamic shared memory: payload tile, next pointers, list heads, column of every key
// (+ per bucket: its length in the current tile and its entry of the skew plan)
__host__ __device__ inline size_t chain_fixed_smem_bytes(int head_cap, int total_dom, int buckets) {
  return (size_t)head_cap * 4 + (size_t)((total_dom + 15) & ~15) + (size_t)buckets * 4 + (size_t)((buckets + 7) & ~7) * 2 + 16;
}
__host__ __device__ inline size_t chain_smem_bytes(int n, int m, int head_cap, int total_dom, int buckets, int tile_rows) {
  return (size_t)tile_rows * (n ? chain_quad_stride(n) : 0) * 16 + (size_t)m * tile_rows * 2 +
         chain_fixed_smem_bytes(head_cap, total_dom, buckets);
}

template <int N>
__global__ void __launch_bounds__(kChainThreads, kChainCtasPerSm) chain_sum_kernel(const __grid_constant__ ChainArgs a) {
  extern __shared__ float4 chain_smem[];
  constexpr int Q = chain_quads(N), QS = N ? chain_quad_stride(N) : 0, QT = Q ? Q : 1;  // QT lanes per list
  const int T = a.tile_rows, m = a.m, D = a.n_groups * a.total_dom, tid = threadIdx.x;
  const int cap = a.head_cap;
  float4 *pay = chain_smem;                                                               // [T][QS]
  unsigned short *nxt = reinterpret_cast<unsigned short *>(pay + (size_t)T * QS);          // [m][T]
  unsigned *head = reinterpret_cast<unsigned *>(nxt + (size_t)m * T);                      // [cap]
  unsigned char *col_of = reinterpret_cast<unsigned char *>(head + cap);                   // [total_dom]
  unsigned *blen = reinterpret_cast<unsigned *>(col_of + ((a.total_dom + 15) & ~15));      // [D] bucket lengths of this tile
  unsigned short *hplan = reinterpret_cast<unsigned short *>(blen + D);  // [D] skew plan: head base (12 bits) | shift << 12
  float *slab = a.slab + (size_t)blockIdx.x * D * (4 * Q);
  unsigned *cslab = a.cnt_slab + (size_t)blockIdx.x * D;
  __shared__ unsigned plan_scan[kChainThreads / 32 + 1];
  // the static plan: the same shift for every bucket, head = (bucket << shift) + (row & (S - 1)), no lookups
  bool uniform = true;
  int ushift = a.sub_shift, heads = D << a.sub_shift;

  for (int c = 0; c < m; c++)
    for (int s = a.cat_off[c] + tid; s < a.cat_off[c + 1]; s += kChainThreads) col_of[s] = (unsigned char)c;

  const unsigned long long n_tiles = (a.n_rows + T - 1) / T;
  int since_fold = 0;
  for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const unsigned long long lo = tile * T;
    const int cnt = (int)min((unsigned long long)T, a.n_rows - lo);
    // (skew plan: a head's upper half-word is its bucket, written when the plan is made)
    for (int i = tid; i < heads; i += kChainThreads) head[i] = uniform ? kChainEnd : ((head[i] & 0xFFFF0000u) | kChainEnd);
    for (int i = tid; i < D; i += kChainThreads) blen[i] = 0;
    __syncthreads();
    // ---- 1. payload rows, list pushes, packed slots
    const bool grouped = a.cols.group != nullptr;
    const int cnt16 = (a.packed && grouped) ? ((cnt + 15) & ~15) : cnt;  // a grouped scan pads its packed columns to 16 rows
    for (int row = tid; row < cnt16; row += kChainThreads) {
      const unsigned long long r = lo + row;
      if (row >= cnt) {  // pad row of the packed columns: not live
        for (int c = 0; c < m + 2; c++) a.packed[c * a.packed_stride + r] = 0;
        continue;
      }
      int g = 0;
      bool live = true;
      if (grouped) {
        g = a.cols.group[r];
        if (g < 0 || g >= a.n_groups) {  // < 0: filtered row
          if (g > 0) atomicExch(a.err, 2);
          live = false;
          g = 0;
        }
      }
      if (!live && !a.packed) continue;
      if constexpr (N > 0) {
        if (live) {
          float v[4 * QT];
#pragma unroll
          for (int k = 0; k < N; k++) v[k] = a.cols.num[k][r];
#pragma unroll
          for (int k = N; k < 4 * Q; k++) v[k] = 0.f;
#pragma unroll
          for (int q = 0; q < Q; q++) pay[(size_t)row * QS + q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
      }
      const int gbase = g * a.total_dom;
      bool bad = false;
      for (int c0 = 0; c0 < m; c0 += 4) {  // keys four columns at a time: the loads are independent
        unsigned s[4];
#pragma unroll
        for (int e = 0; e < 4; e++) s[e] = c0 + e < m ? (unsigned)(a.cols.cat[c0 + e][r] - a.lo[c0 + e]) : 0u;
#pragma unroll
        for (int e = 0; e < 4; e++)
          if (c0 + e < m) {
            const int c = c0 + e;
            const bool ok = s[e] < (unsigned)a.dom[c];
            if (ok && live) {
              const int bucket = gbase + a.cat_off[c] + (int)s[e];
              int h;
              unsigned tag = (unsigned)row;
              if (uniform) {
                h = (bucket << ushift) + (row & ((1 << ushift) - 1));
              } else {
                const unsigned pl = hplan[bucket];
                h = (int)(pl & 0xFFFu) + (row & ((1 << (pl >> 12)) - 1));
                tag |= (unsigned)bucket << 16;
              }
              const unsigned prev = atomicExch(&head[h], tag);
              nxt[(size_t)c * T + row] = (unsigned short)prev;
            }
            // a filtered row keeps its real slots: its zero increments then spread over the pair tables like live rows
            if (a.packed) a.packed[c * a.packed_stride + r] = ok ? (unsigned char)s[e] : (unsigned char)0;
            bad |= !ok && live;
          }
      }
      if (a.packed && grouped) {
        a.packed[m * a.packed_stride + r] = (unsigned char)g;
        a.packed[(m + 1) * a.packed_stride + r] = live ? (unsigned char)1 : (unsigned char)0;
      }
      if (bad) atomicExch(a.err, 1);  // a key outside the declared domain: the scan reports CFB_ERR_DOMAIN
    }
    __syncthreads();
    // the next tile of this CTA is pulled into L2 while the lists are walked: pass 1 is bound by the latency of its
    // global loads (a few dependent batches per row), and an L2 hit costs a third of an HBM access
    if (tile + gridDim.x < n_tiles) {
      const unsigned long long nlo = (tile + gridDim.x) * T;
      const int ncnt = (int)min((unsigned long long)T, a.n_rows - nlo);
      const int lines = (ncnt + 31) / 32, n_cols = N + m + (a.cols.group ? 1 : 0);
      for (int i = tid; i < lines * n_cols; i += kChainThreads) {
        const int col = i / lines, line = i - col * lines;
        const void *p = col < N ? (const void *)(a.cols.num[col] + nlo + 32 * line)
                                : (col < N + m ? (const void *)(a.cols.cat[col - N] + nlo + 32 * line)
                                               : (const void *)(a.cols.group + nlo + 32 * line));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
      }
    }
    // ---- 2. list walks: QT adjacent lanes per list, one payload quad each; sums and counts to the CTA's slab.
    //         A lane follows TWO lists at a time: a list step is a dependent pair of shared-memory loads (next pointer,
    //         payload), and one list per lane leaves the load pipe waiting on that latency.
    const int n_tasks = heads * QT;
    for (int task = tid; task < n_tasks; task += 2 * kChainThreads) {
      const int task_b = task + kChainThreads;
      const int list_a = task / QT, q_a = task - list_a * QT;
      const int list_b = task_b < n_tasks ? task_b / QT : list_a, q_b = task_b - (task_b / QT) * QT;
      const unsigned head_a = head[list_a], head_b = task_b < n_tasks ? head[list_b] : kChainEnd;
      unsigned row_a = head_a & 0xFFFFu, row_b = head_b & 0xFFFFu;
      if (row_a == kChainEnd && row_b == kChainEnd) continue;
      const int b_a = uniform ? list_a >> ushift : (int)(head_a >> 16);  // bucket = (slot, column, key)
      const int b_b = uniform ? list_b >> ushift : (int)(head_b >> 16);
      const unsigned short *nx_a = nxt + (size_t)col_of[b_a % a.total_dom] * T;  // the column's next pointers
      const unsigned short *nx_b = nxt + (size_t)col_of[b_b % a.total_dom] * T;
      float4 acc_a = make_float4(0.f, 0.f, 0.f, 0.f), acc_b = acc_a;
      unsigned n_a = 0, n_b = 0;
      while (row_a != kChainEnd && row_b != kChainEnd) {
        const unsigned nrow_a = nx_a[row_a], nrow_b = nx_b[row_b];
        if constexpr (N > 0) {
          const float4 wa = pay[(size_t)row_a * QS + q_a], wb = pay[(size_t)row_b * QS + q_b];
          acc_a.x += wa.x, acc_a.y += wa.y, acc_a.z += wa.z, acc_a.w += wa.w;
          acc_b.x += wb.x, acc_b.y += wb.y, acc_b.z += wb.z, acc_b.w += wb.w;
        }
        n_a++, n_b++;
        row_a = nrow_a, row_b = nrow_b;
      }
      while (row_a != kChainEnd) {
        const unsigned nrow = nx_a[row_a];
        if constexpr (N > 0) {
          const float4 w = pay[(size_t)row_a * QS + q_a];
          acc_a.x += w.x, acc_a.y += w.y, acc_a.z += w.z, acc_a.w += w.w;
        }
        n_a++;
        row_a = nrow;
      }
      while (row_b != kChainEnd) {
        const unsigned nrow = nx_b[row_b];
        if constexpr (N > 0) {
          const float4 w = pay[(size_t)row_b * QS + q_b];
          acc_b.x += w.x, acc_b.y += w.y, acc_b.z += w.z, acc_b.w += w.w;
        }
        n_b++;
        row_b = nrow;
      }
      if (n_a) {
        if constexpr (N > 0) red_v4(slab + (size_t)b_a * (4 * Q) + 4 * q_a, acc_a.x, acc_a.y, acc_a.z, acc_a.w);
        if (q_a == 0) atomicAdd(blen + b_a, n_a);
      }
      if (n_b) {
        if constexpr (N > 0) red_v4(slab + (size_t)b_b * (4 * Q) + 4 * q_b, acc_b.x, acc_b.y, acc_b.z, acc_b.w);
        if (q_b == 0) atomicAdd(blen + b_b, n_b);
      }
    }
    __syncthreads();
    // ---- 2b. key counts to the count slab; the sub-list plan of the next tile from this tile's bucket lengths
    {
      constexpr int kTargetLen = 32, kMaxShift = 8;
      auto want = [&](unsigned len, int target) {
        int sh = 0;
        while (sh < kMaxShift && (len + target - 1) / target > (1u << sh)) sh++;
        return sh;
      };
      // the static plan is good enough while its longest sub-list stays within a few times the average work of a
      // lane slot (two lists per lane in flight): imbalance only costs when one walk outlasts everything else
      const unsigned slot_work = (unsigned)((long long)cnt * m * QT / (2 * kChainThreads)) + 1;
      int hot = 0;
      for (int b = tid; b < D; b += kChainThreads) {
        const unsigned len = blen[b];
        if (len) atomicAdd(cslab + b, len);
        hot |= (len >> a.sub_shift) > 4 * slot_work;
      }
      const bool skewed = __syncthreads_or(hot) != 0 && a.adaptive;
      if (!skewed) {
        uniform = true;
        ushift = a.sub_shift;
        heads = D << a.sub_shift;
      } else {
        // per bucket 2^shift sub-lists of ~target rows; the target doubles until the heads fit
        const int per = (D + kChainThreads - 1) / kChainThreads, b0 = tid * per, b1 = min(D, b0 + per);
        int target = kTargetLen;
        unsigned mine = 0, total = 0, before = 0;
        for (;; target *= 2) {
          mine = 0;
          for (int b = b0; b < b1; b++) mine += 1u << want(blen[b], target);
          unsigned incl = mine;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((tid & 31) >= o) incl += t;
          }
          __syncthreads();
          if ((tid & 31) == 31) plan_scan[tid >> 5] = incl;
          __syncthreads();
          before = incl - mine;
          total = 0;
          for (int w = 0; w < kChainThreads / 32; w++) {
            const unsigned v = plan_scan[w];
            if (w < (tid >> 5)) before += v;
            total += v;
          }
          if ((int)total <= cap || target >= (1 << 20)) break;
        }
        unsigned base = before;
        for (int b = b0; b < b1; b++) {
          const int sh = want(blen[b], target);
          hplan[b] = (unsigned short)(base | ((unsigned)sh << 12));
          for (unsigned j = 0; j < (1u << sh); j++) head[base + j] = ((unsigned)b << 16) | kChainEnd;
          base += 1u << sh;
        }
        uniform = false;
        heads = (int)total;
      }
    }
    // ---- 3. fold the slab into the fp64 / u64 state
    if (++since_fold >= a.fold_tiles || tile + gridDim.x >= n_tiles) {
      since_fold = 0;
      __threadfence();
      __syncthreads();
      for (int i = tid; i < D * 4 * Q; i += kChainThreads) {
        const float v = __ldcg(slab + i);
        if (v == 0.f) continue;
        __stcg(slab + i, 0.f);
        const int b = i / (4 * QT), j = i % (4 * QT), g = b / a.total_dom, key = b % a.total_dom;
        if (j < N) atomicAdd(a.f64 + g * a.F + a.numcat_base + (long long)j * a.total_dom + key, (double)v);
      }
      for (int b = tid; b < D; b += kChainThreads) {
        const unsigned v = __ldcg(cslab + b);
        if (!v) continue;
        __stcg(cslab + b, 0u);
        red_u64(a.u64 + (b / a.total_dom) * a.U + 1 + b % a.total_dom, v);
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------
// Pair counts from the packed slots: every thread takes 16 consecutive rows of a table's two columns as two
// 128-bit loads (16 one-byte slots each); a pair update is two byte extracts (PRMT), two multiply-adds for the cell
// address and one red.shared -- 5 instructions against the ~23 of role_scan_kernel, which re-derives and re-validates
// slots from the int32 columns for every table.  Tables, roles and replicas as in role_kernels.cuh (RolePlan).
// GROUPED scans (GROUP BY slots / row filter) read two more packed columns: the slot selects the sub-table, the live
// byte IS the increment (0 for a filtered row: no predicate, no branch; such rows keep their real slots, so their
// zero increments spread over the table like everyone else's).
struct PackedPairArgs {
  const unsigned char *packed;
  unsigned long long stride;
  unsigned long long n_rows;  // GROUPED: a multiple of 16 (pad rows are not live); else exact, the tail is done by row
  int chunk_rows;             // multiple of 16
  int pair_fold_chunks;       // fold the tables every this many chunks of a CTA (16-bit cells: <= kRoleFoldRows16 rows)
  int n_reps;                 // replicas per role: gridDim.x = n_roles * n_reps
  int m, n_groups;
  long long U, pair_base;
  unsigned long long *u64;
  RolePlan plan;
};

__device__ __forceinline__ unsigned packed_byte(unsigned word, int b) { return __byte_perm(word, 0u, 0x4440u + b); }

template <int BITS, bool GROUPED>
__global__ void __launch_bounds__(kRoleThreads, 1) pair_packed_kernel(const __grid_constant__ PackedPairArgs a) {
  extern __shared__ unsigned pair_smem[];  // [plan.words[role]] pair tables of this CTA's role
  const int n_roles = a.plan.n_roles;
  const int role = blockIdx.x % n_roles, rep = blockIdx.x / n_roles;
  const int nt = a.plan.n_tables[role], words = a.plan.words[role];
  for (int i = threadIdx.x; i < words; i += kRoleThreads) pair_smem[i] = 0;
  __syncthreads();
  unsigned long long *pairs = a.u64 + a.pair_base;
  const unsigned smem_base = (unsigned)__cvta_generic_to_shared(pair_smem);
  const unsigned long long rows16 = a.n_rows & ~15ull;  // rows covered by whole 16-row groups
  const unsigned long long n_chunks = (rows16 + a.chunk_rows - 1) / a.chunk_rows;
  const unsigned char *cg = a.packed + (unsigned long long)a.m * a.stride, *cv = cg + a.stride;
  int since_fold = 0;
  auto update = [&](unsigned tbl, unsigned dom_l4, unsigned gcells, unsigned sk, unsigned sl, unsigned g, unsigned live) {
    if constexpr (BITS == 32) {
      unsigned addr = sl * 4u + tbl;
      addr = sk * dom_l4 + addr;
      if constexpr (GROUPED) {
        addr = g * (4u * gcells) + addr;
        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(live) : "memory");
      } else {
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
      }
    } else {
      unsigned cell = sk * (dom_l4 >> 2) + sl;
      if constexpr (GROUPED) cell += g * gcells;
      const unsigned addr = tbl + 4u * (cell >> 1);
      const unsigned inc = (GROUPED ? live : 1u) << ((cell & 1u) * 16);
      asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(inc) : "memory");
    }
  };
  for (unsigned long long ch = rep; ch < n_chunks || (ch == rep && n_chunks == 0); ch += a.n_reps) {
    const unsigned long long lo = ch * a.chunk_rows, hi = min(rows16, lo + (unsigned long long)a.chunk_rows);
    const bool last_of_cta = ch + a.n_reps >= n_chunks;
    for (int t = 0; t < nt; t++) {
      const RoleTable d = a.plan.tbl[role][t];
      const unsigned char *ck = a.packed + d.k * a.stride, *cl = a.packed + d.l * a.stride;
      const unsigned dom_l4 = 4u * (unsigned)d.dom_l, tbl = smem_base + 4u * (unsigned)d.word_off;
      const unsigned gcells = (unsigned)d.gwords * (BITS == 32 ? 1u : 2u);
      for (unsigned long long r = lo + 16ull * threadIdx.x; r < hi; r += 16ull * kRoleThreads) {
        const uint4 kv = *reinterpret_cast<const uint4 *>(ck + r), lv = *reinterpret_cast<const uint4 *>(cl + r);
        uint4 gv = make_uint4(0u, 0u, 0u, 0u), vv = gv;
        if constexpr (GROUPED) {
          if (a.n_groups > 1) gv = *reinterpret_cast<const uint4 *>(cg + r);
          vv = *reinterpret_cast<const uint4 *>(cv + r);
        }
        const unsigned kw[4] = {kv.x, kv.y, kv.z, kv.w}, lw[4] = {lv.x, lv.y, lv.z, lv.w};
        const unsigned gw[4] = {gv.x, gv.y, gv.z, gv.w}, vw[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
        for (int w = 0; w < 4; w++)
#pragma unroll
          for (int b = 0; b < 4; b++)
            update(tbl, dom_l4, gcells, packed_byte(kw[w], b), packed_byte(lw[w], b), GROUPED ? packed_byte(gw[w], b) : 0u,
                   GROUPED ? packed_byte(vw[w], b) : 1u);
      }
      // the rows past the last whole 16-row group (ungrouped scans only), one thread per row, once per role
      if (!GROUPED && rep == 0 && last_of_cta) {
        const unsigned long long r = rows16 + threadIdx.x;
        if (r < a.n_rows) update(tbl, dom_l4, gcells, ck[r], cl[r], 0u, 1u);
      }
    }
    __syncthreads();
    // fold the pair tables into the u64 state and zero them
    const bool fold = ++since_fold >= a.pair_fold_chunks || last_of_cta;
    if (!fold) continue;
    since_fold = 0;
    for (int tg = 0; tg < nt * a.n_groups; tg++) {
      const RoleTable &d = a.plan.tbl[role][tg / a.n_groups];
      const int g = tg % a.n_groups;
      unsigned long long *dst = pairs + g * a.U + d.state_off;
      unsigned *src = pair_smem + d.word_off + g * d.gwords;
      if constexpr (BITS == 32) {
        for (int i = threadIdx.x; i < d.cells; i += kRoleThreads) {
          const unsigned v = src[i];
          if (v) {
            src[i] = 0;
            red_u64(dst + i, v);
          }
        }
      } else {
        for (int i = threadIdx.x; i < (d.cells + 1) / 2; i += kRoleThreads) {
          const unsigned v = src[i];
          if (v) {
            src[i] = 0;
            if (v & 0xffffu) red_u64(dst + 2 * i, v & 0xffffu);
            if (v >> 16) red_u64(dst + 2 * i + 1, v >> 16);
          }
        }
      }
    }
    __syncthreads();
    if (n_chunks == 0) break;
  }
}

}  // namespace cfb
