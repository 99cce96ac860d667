// role_kernels.cuh -- kernel family K3 (categorical aggregates): (key1,key2) pair counts in shared-memory tables.
//
// Replaces the per-row std::map updates of Triple::SumNoLift (sum_no_lift.cpp:195-214, and :158-189 when
// bucket_sum_kernel does not apply) for the common case of dense, small categorical domains (e.g. BASELINE
// config 3: 10 columns of domain 100).
//
// The categorical part is bound by the NUMBER of scattered updates a row makes (C3: 45 pair counts + 10 x
// [count, 10 sums]); csrc/micro/cat_probe.cu (profiles/r01_cat_probe.txt) measured the units that can execute
// them on the B200: L2 reductions ~200 G ops/s chip-wide whatever their width, shared-memory atomics 2-4x that;
// match.any de-duplication + plain LDS/STS and warp-private plain read-modify-write tables are slower than
// ATOMS; L2 reductions and ATOMS issued together do not overlap.  So:
//   * pair counts go to shared-memory tables with ATOMS.  All pair tables do not fit one SM (C3: 45 x 10^4
//     cells), so they are cut into ROLES (RolePlan, built on the host, passed by value in the kernel
//     parameters): a CTA of role r holds that role's tables (one sub-table per GROUP BY slot) and scans ALL
//     rows of its chunks for them; roles x replicas CTAs cover the grid.  Cells are 32-bit, or 16-bit packed
//     two per word when 32-bit cells would need too many roles (16-bit tables are folded at least every
//     65 024 rows, so a cell cannot overflow);
//   * a chunk is scanned once per table of the role, with the table's descriptor in registers (a loop over the
//     tables inside the row loop is instruction bound); every thread takes 4 consecutive rows per step, each key
//     column as one 128-bit load (only wide loads keep enough bytes in flight to hide the L2 / HBM latency), and
//     the increments are predicated red.shared, not branches around atomics;
//   * the per-key payloads [count, x_0..x_{n-1}] are normally done by bucket_sum_kernel (no float atomics).
//     When that kernel does not apply (sum of domains > 4096) they go out from here as 128-bit vector
//     reductions into per-CTA fp32 slabs in L2, every 4096-row tile by ONE role (rotating), all columns;
//   * at the end of a chunk the CTA folds its tables / slab into the u64 / fp64 state.
// Shapes this kernel does not take (hashed pairs, tables larger than shared memory, more roles than SMs / 8)
// stay on slab_scan_kernel.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "slab_kernels.cuh"
#include "state_layout.h"

namespace cfb {

constexpr int kRoleThreads = 1024;
constexpr int kRoleMaxRoles = 12;
constexpr int kRoleMaxTables = 24;       // pair tables per role
constexpr int kRoleMaxChunkRows = 32512;  // an fp32 slab cell is folded into fp64 after at most this many rows
constexpr int kRoleFoldRows16 = 65024;    // <= 65535: rows a 16-bit cell can take between two folds

struct RoleTable {
  int k, l;             // the pair (k < l)
  int dom_l;            // cell = slot_k * dom_l + slot_l
  int cells;            // dom_k * dom_l (per GROUP BY slot)
  int gwords;           // 32-bit words of one slot's cells; the table holds n_groups of them
  int word_off;         // first 32-bit word of the table in the role's shared memory
  int state_off;        // Layout::pair_off[k*m+l] (dense pair regions are < 2 GB: kDensePairBytes)
};
// Passed BY VALUE in the kernel parameters: every field a thread reads in the row loop is uniform over the
// CTA, so it comes through the constant bank (LDC / uniform registers) and costs no load/store-unit slot.
struct RolePlan {
  int n_roles, bits;  // bits = 16 or 32
  int n_tables[kRoleMaxRoles];
  int words[kRoleMaxRoles];  // shared-memory words of the role's tables
  RoleTable tbl[kRoleMaxRoles][kRoleMaxTables];
};

struct RoleArgs {
  ScanCols cols;
  unsigned long long n_rows;
  int chunk_rows;        // <= kRoleMaxChunkRows
  int pair_fold_chunks;  // fold the pair tables every this many chunks of a CTA (16-bit cells: <= kRoleFoldRows16 rows)
  int n_reps;            // replicas per role: gridDim.x = n_roles * n_reps
  int m, n_groups;
  long long U;  // per-slot stride of the u64 state
  int skip;  // bit 0: no pair counts (measurement only), bit 1: no per-key payloads (bucket_sum_kernel does them)
  int lo[kMaxCat], dom[kMaxCat], cat_off[kMaxCat + 1];  // of the Layout
  long long total_dom, numcat_base, pair_base;
  int n_sub;    // fp32 slabs per CTA (1, 2 or 4): threads are spread over them to thin out same-address reductions
  float *slab;  // [gridDim.x * n_sub][pad4(total_dom * P)], all zero on entry and on exit
  double *f64;
  unsigned long long *u64;
  int *err;
  RolePlan plan;
};

// 4 consecutive rows of a column: one 128-bit load (columns are 16-byte aligned and r is a multiple of 4),
// scalar loads at the tail of the table.
template <class T4, class T>
__device__ __forceinline__ T4 load_rows4(const T *col, unsigned long long r, unsigned long long hi) {
  if (r + 4 <= hi) return *reinterpret_cast<const T4 *>(col + r);
  T4 v;
  v.x = col[r];
  v.y = r + 1 < hi ? col[r + 1] : T(0);
  v.z = r + 2 < hi ? col[r + 2] : T(0);
  v.w = T(0);
  return v;
}

// Pair counts of ONE table over the rows [lo, hi) of a chunk: the table's descriptor lives in registers for the
// whole pass (a loop over the tables inside the row loop re-reads it from the constant bank per row step and is
// instruction bound).  4 consecutive rows per thread and step, keys as 128-bit loads, predicated red.shared.
template <int BITS, bool TAIL>
__device__ __forceinline__ unsigned role_table_pass(const RoleArgs &a, const RoleTable &d, unsigned smem_base,
                                                    unsigned long long lo, unsigned long long hi) {
  const int *ck = a.cols.cat[d.k], *cl = a.cols.cat[d.l], *grp = a.cols.group;
  const int lo_k = a.lo[d.k], lo_l = a.lo[d.l], G = a.n_groups;
  const unsigned dom_k = (unsigned)a.dom[d.k], dom_l = (unsigned)d.dom_l;
  const unsigned tbl = smem_base + 4u * (unsigned)d.word_off, gcells = (unsigned)d.gwords * (BITS == 32 ? 1u : 2u);
  unsigned bad = 0;
  for (unsigned long long r = lo + 4ull * threadIdx.x; r < hi; r += 4ull * kRoleThreads) {
    unsigned on = 1u | (!TAIL || r + 1 < hi ? 2u : 0u) | (!TAIL || r + 2 < hi ? 4u : 0u) | (!TAIL || r + 3 < hi ? 8u : 0u);
    int4 kv, lv, gv = make_int4(0, 0, 0, 0);
    if constexpr (TAIL) {
      kv = load_rows4<int4>(ck, r, hi);
      lv = load_rows4<int4>(cl, r, hi);
      if (grp) gv = load_rows4<int4>(grp, r, hi);
    } else {
      kv = *reinterpret_cast<const int4 *>(ck + r);
      lv = *reinterpret_cast<const int4 *>(cl + r);
      if (grp) gv = *reinterpret_cast<const int4 *>(grp + r);
    }
    const int g[4] = {gv.x, gv.y, gv.z, gv.w};
    const unsigned sk[4] = {(unsigned)(kv.x - lo_k), (unsigned)(kv.y - lo_k), (unsigned)(kv.z - lo_k), (unsigned)(kv.w - lo_k)};
    const unsigned sl[4] = {(unsigned)(lv.x - lo_l), (unsigned)(lv.y - lo_l), (unsigned)(lv.z - lo_l), (unsigned)(lv.w - lo_l)};
    unsigned in = 0, hit = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      in |= ((unsigned)g[i] < (unsigned)G ? 1u : 0u) << i;  // slot < 0: filtered row (>= G is reported by the payload pass / bucket_sum_kernel)
      hit |= (sk[i] < dom_k && sl[i] < dom_l ? 1u : 0u) << i;
    }
    on &= in;
    bad |= on & ~hit;
    hit &= on;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      unsigned cell = sk[i] * dom_l + sl[i];
      if (G > 1) cell += (unsigned)g[i] * gcells;
      const unsigned addr = tbl + (BITS == 32 ? 4u * cell : 4u * (cell >> 1));
      const unsigned inc = BITS == 32 ? 1u : 1u << ((cell & 1u) * 16);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p red.shared.add.u32 [%0], %1;\n\t}" ::"r"(addr), "r"(inc), "r"(hit & (1u << i)) : "memory");
    }
  }
  return bad;
}

// Payload step of a thread (only when bucket_sum_kernel does not apply): the per-key payloads of 4 consecutive rows
// [r, r+4) when this tile is this role's; also reports GROUP BY slots >= n_groups.  TAIL: the rows may run past `hi`.
template <int N, int BITS, bool TAIL>
__device__ __forceinline__ void role_step(const RoleArgs &a, int role, int nt, unsigned *role_smem, float *slab,
                                          unsigned long long r, unsigned long long hi, bool sums_mine) {
  constexpr int P = pad4(1 + N);
  auto rows4 = [&](const int *col) {
    if constexpr (TAIL) return load_rows4<int4>(col, r, hi);
    else return *reinterpret_cast<const int4 *>(col + r);
  };
  // bit i of a mask = row r + i
  unsigned on = 1u | (!TAIL || r + 1 < hi ? 2u : 0u) | (!TAIL || r + 2 < hi ? 4u : 0u) | (!TAIL || r + 3 < hi ? 8u : 0u);
  int gv[4] = {0, 0, 0, 0};  // GROUP BY slot of the row
  unsigned bad = 0;
  if (a.cols.group) {
    const int4 g = rows4(a.cols.group);
    gv[0] = g.x, gv[1] = g.y, gv[2] = g.z, gv[3] = g.w;
    unsigned in = 0, over = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      in |= ((unsigned)gv[i] < (unsigned)a.n_groups ? 1u : 0u) << i;  // < 0: filtered row
      over |= (gv[i] >= a.n_groups ? 1u : 0u) << i;
    }
    if (on & over) atomicExch(a.err, 2);
    on &= in;
  }
  // per-key payload [1, x_0..x_{N-1}] of all columns when this tile is this role's: L2 vector reductions,
  // one quad of the payload at a time (4 rows x 4 values in registers)
  if (sums_mine && !(a.skip & 2)) {
#pragma unroll
    for (int q = 0; q < P / 4; q++) {
      float4 x[4];  // x[e] = payload element 4q+e of the 4 rows
#pragma unroll
      for (int e = 0; e < 4; e++) {
        constexpr int kNone = -1;
        const int col = 4 * q + e - 1;  // payload element 0 is the count
        if (col == kNone) x[e] = make_float4(1.f, 1.f, 1.f, 1.f);
        else if (col < N) x[e] = load_rows4<float4>(a.cols.num[col < N ? col : 0], r, hi);
        else x[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      for (int c = 0; c < a.m; c++) {
        const int4 v = rows4(a.cols.cat[c]);
        const int lo_c = a.lo[c];
        const unsigned sc[4] = {(unsigned)(v.x - lo_c), (unsigned)(v.y - lo_c), (unsigned)(v.z - lo_c), (unsigned)(v.w - lo_c)};
        const float xr[4][4] = {{x[0].x, x[1].x, x[2].x, x[3].x}, {x[0].y, x[1].y, x[2].y, x[3].y},
                                {x[0].z, x[1].z, x[2].z, x[3].z}, {x[0].w, x[1].w, x[2].w, x[3].w}};
#pragma unroll
        for (int i = 0; i < 4; i++) {
          if (!(on & (1u << i))) continue;
          if (sc[i] >= (unsigned)a.dom[c]) {
            bad |= 1u;
            continue;
          }
          red_v4(slab + (size_t)(a.cat_off[c] + (int)sc[i]) * P + 4 * q, xr[i][0], xr[i][1], xr[i][2], xr[i][3]);
        }
      }
    }
  }
  if (bad) atomicExch(a.err, 1);  // a key outside the declared domain: the scan reports CFB_ERR_DOMAIN
}

template <int N, int BITS>
__global__ void __launch_bounds__(kRoleThreads, 1) role_scan_kernel(const __grid_constant__ RoleArgs a) {
  extern __shared__ unsigned role_smem[];  // [plan.words[role]] pair tables of this CTA's role
  constexpr int P = pad4(1 + N);
  const int n_roles = a.plan.n_roles;
  const int role = blockIdx.x % n_roles, rep = blockIdx.x / n_roles;
  const int nt = a.plan.n_tables[role], words = a.plan.words[role], m = a.m;
  for (int i = threadIdx.x; i < words; i += kRoleThreads) role_smem[i] = 0;
  __syncthreads();
  const long long slab_floats = (a.total_dom * P + 3) & ~3ll;
  float *cta_slabs = a.slab + (size_t)blockIdx.x * a.n_sub * slab_floats;
  float *slab = cta_slabs + (size_t)(threadIdx.x * a.n_sub / kRoleThreads) * slab_floats;
  unsigned long long *pairs = a.u64 + a.pair_base;

  const unsigned long long n_chunks = (a.n_rows + a.chunk_rows - 1) / a.chunk_rows;
  int since_fold = 0, slab_rows = 0;
  for (unsigned long long ch = rep; ch < n_chunks; ch += a.n_reps) {
    const unsigned long long lo = ch * a.chunk_rows, hi = min(a.n_rows, lo + (unsigned long long)a.chunk_rows);
    // Every role sees every row; the per-key payloads of a 4096-row tile are issued by ONE role (all
    // columns: the reductions of a CTA then spread over all keys of all columns, which matters because
    // L2 serialises reductions to one address), rotating over the roles tile by tile.
    int tile = (int)(ch % n_roles);
    // the last chunk of a table whose row count is not a multiple of 4 takes the variants with tail checks
    const bool tail = (hi - lo) % 4 != 0;
    const unsigned smem_base = (unsigned)__cvta_generic_to_shared(role_smem);
    if (!(a.skip & 1)) {
      unsigned bad = 0;
      for (int t = 0; t < nt; t++) {
        const RoleTable d = a.plan.tbl[role][t];
        bad |= tail ? role_table_pass<BITS, true>(a, d, smem_base, lo, hi) : role_table_pass<BITS, false>(a, d, smem_base, lo, hi);
      }
      if (bad) atomicExch(a.err, 1);  // a key outside the declared domain: the scan reports CFB_ERR_DOMAIN
    }
    if (!(a.skip & 2)) {  // payloads, when bucket_sum_kernel does not do them
      if (!tail) {
        for (unsigned long long r = lo + 4ull * threadIdx.x; r < hi; r += 4ull * kRoleThreads) {
          role_step<N, BITS, false>(a, role, nt, role_smem, slab, r, hi, tile == role);
          tile = tile + 1 == n_roles ? 0 : tile + 1;
        }
      } else {
        for (unsigned long long r = lo + 4ull * threadIdx.x; r < hi; r += 4ull * kRoleThreads) {
          role_step<N, BITS, true>(a, role, nt, role_smem, slab, r, hi, tile == role);
          tile = tile + 1 == n_roles ? 0 : tile + 1;
        }
      }
    }
    __threadfence();
    __syncthreads();
    // fold the pair tables into the u64 state and zero them
    const bool fold_pairs = ++since_fold >= a.pair_fold_chunks || ch + a.n_reps >= n_chunks;
    if (fold_pairs) since_fold = 0;
    for (int tg = 0; fold_pairs && tg < nt * a.n_groups; tg++) {
      const RoleTable &d = a.plan.tbl[role][tg / a.n_groups];
      const int g = tg % a.n_groups;
      unsigned long long *dst = pairs + g * a.U + d.state_off;
      unsigned *src = role_smem + d.word_off + g * d.gwords;
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
    // fold this CTA's slab into the fp64 / u64 state once it has taken kRoleMaxChunkRows rows (bounds every fp32 run)
    for (int t = 0, who = (int)(ch % n_roles); t * 4 * kRoleThreads < (int)(hi - lo); t++, who = who + 1 == n_roles ? 0 : who + 1)
      if (who == role) slab_rows += 4 * kRoleThreads;
    if (!(a.skip & 2) && (slab_rows + 2 * 4 * kRoleThreads > kRoleMaxChunkRows || ch + a.n_reps >= n_chunks)) {
      slab_rows = 0;
      for (long long i = threadIdx.x; i < a.total_dom * P; i += kRoleThreads) {
        float v = 0.f;
        for (int u = 0; u < a.n_sub; u++) {
          const float w = __ldcg(cta_slabs + u * slab_floats + i);
          if (w != 0.f) {
            __stcg(cta_slabs + u * slab_floats + i, 0.f);
            v += w;
          }
        }
        if (v == 0.f) continue;
        const long long tkey = i / P;
        const int j = (int)(i % P);
        if (j == 0)
          red_u64(a.u64 + 1 + tkey, (unsigned long long)v);
        else if (j <= N)
          atomicAdd(a.f64 + a.numcat_base + (long long)(j - 1) * a.total_dom + tkey, (double)v);
      }
    }
    __threadfence();
    __syncthreads();
  }
}

}  // namespace cfb
