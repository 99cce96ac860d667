// scatter_kernels.cuh -- the small kernels around the scans: the scalar-atomic fallback scan, key
// range pre-pass, combine / re-layout of dense states (remap_add), draining of hashed pair counts,
// the sum_triple / sum_nb_agg kernels over lifted triples, and the synthetic-data generators.
// (The main categorical / GROUP BY kernels are slab_kernels.cuh and group_kernel.cuh.)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "pair_hash.cuh"
#include "state_layout.h"

namespace cfb {

__device__ __forceinline__ void add_u64(unsigned long long *p, unsigned long long v) { atomicAdd(p, v); }

// ---------------------------------------------------------------------------------------
// Generic scan: one thread per row, every update a scalar atomic on the context state.  The
// fallback behind slab_scan_kernel (slab_kernels.cuh) for shapes whose per-CTA slab would be too
// large (very many GROUP BY slots x wide domains).  do_numeric = also accumulate N / lin / quad.
__global__ void __launch_bounds__(256)
    generic_scan_kernel(const ScanCols cols, const Layout *__restrict__ lay_g, unsigned long long n_rows,
                        int do_numeric, double *__restrict__ f64, unsigned long long *__restrict__ u64,
                        int *__restrict__ err, const PairHash hash) {
  __shared__ Layout lay;
  {
    const int *src = reinterpret_cast<const int *>(lay_g);
    int *dst = reinterpret_cast<int *>(&lay);
    for (int i = threadIdx.x; i < (int)(sizeof(Layout) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const int n = lay.n, m = lay.m;
  const bool triple = lay.kind == 0;
  for (unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows;
       r += (unsigned long long)gridDim.x * blockDim.x) {
    int g = cols.group ? cols.group[r] : 0;
    if (g < 0) continue;  // filtered row
    if (g >= lay.n_groups) {
      atomicExch(err, 2);
      continue;
    }
    double *F = f64 + (long long)g * lay.F;
    unsigned long long *U = u64 + (long long)g * lay.U;
    float x[32];
    for (int k = 0; k < n; k++) x[k] = cols.num[k][r];
    if (do_numeric) {
      add_u64(U, 1ull);
      for (int k = 0; k < n; k++) atomicAdd(F + k, (double)x[k]);
      if (triple) {
        int p = n;
        for (int i = 0; i < n; i++)
          for (int j = i; j < n; j++, p++) atomicAdd(F + p, (double)x[i] * (double)x[j]);
      } else {
        for (int i = 0; i < n; i++) atomicAdd(F + n + i, (double)x[i] * (double)x[i]);
      }
    }
    int s[32];
    bool ok = true;
    for (int c = 0; c < m; c++) {
      const long long d = (long long)cols.cat[c][r] - (long long)lay.lo[c];
      if (d < 0 || d >= lay.dom[c]) ok = false;
      s[c] = (int)d;
    }
    if (!ok) {
      atomicExch(err, 1);
      continue;
    }
    for (int c = 0; c < m; c++) {
      const long long t = lay.cat_off[c] + s[c];
      add_u64(U + 1 + t, 1ull);
      if (triple)
        for (int i = 0; i < n; i++) atomicAdd(F + lay.numcat_base + (long long)i * lay.total_dom + t, (double)x[i]);
    }
    if (triple)
      for (int k = 0; k < m; k++)
        for (int l = k + 1; l < m; l++) {
          if (!lay.pairs_hashed)
            add_u64(U + lay.pair_base + lay.pair_off[k * m + l] + (long long)s[k] * lay.dom[l] + s[l], 1ull);
          else if (!pair_hash_add(hash, g, pair_key(k * m + l, s[k], s[l]), 1ull))
            atomicExch(err, 3);
        }
  }
}

// ---------------------------------------------------------------------------------------
// N += rows for an ungrouped scan (the count needs no pass over the data).
__global__ void add_rows_kernel(unsigned long long *u64, unsigned long long rows) {
  if (threadIdx.x == 0 && blockIdx.x == 0) u64[0] += rows;
}

// ---------------------------------------------------------------------------------------
// Observed [min,max] of categorical columns (blockIdx.y = column).
__global__ void __launch_bounds__(256)
    cat_minmax_kernel(const ScanCols cols, unsigned long long n_rows, int *__restrict__ lo, int *__restrict__ hi) {
  const int32_t *col = cols.cat[blockIdx.y];
  int mn = INT32_MAX, mx = INT32_MIN;
  for (unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows;
       r += (unsigned long long)gridDim.x * blockDim.x) {
    const int v = col[r];
    mn = min(mn, v);
    mx = max(mx, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0 && mn <= mx) {
    atomicMin(lo + blockIdx.y, mn);
    atomicMax(hi + blockIdx.y, mx);
  }
}

// ---------------------------------------------------------------------------------------
// dst += src where both are dense states with the same (kind, n, m, n_groups) and src's
// categorical domain lies inside dst's.  Used by combine (sum_state.cpp:23-112), by domain
// growth and by import/export of partials.  One thread per src element.
__global__ void __launch_bounds__(256)
    remap_add_kernel(const Layout *__restrict__ dl_g, const Layout *__restrict__ sl_g, double *__restrict__ df,
                     unsigned long long *__restrict__ du, const double *__restrict__ sf,
                     const unsigned long long *__restrict__ su, const PairHash dhash, int *__restrict__ err,
                     const SlotTrans tr, const int *__restrict__ group_map) {
  // group_map: dst GROUP BY slot of every src slot (nullptr = identity, < 0 = skip); must be injective
  __shared__ Layout dl, sl;
  {
    const int *a = reinterpret_cast<const int *>(dl_g), *b = reinterpret_cast<const int *>(sl_g);
    int *da = reinterpret_cast<int *>(&dl), *db = reinterpret_cast<int *>(&sl);
    for (int i = threadIdx.x; i < (int)(sizeof(Layout) / sizeof(int)); i += blockDim.x) {
      da[i] = a[i];
      db[i] = b[i];
    }
  }
  __syncthreads();
  const int n = sl.n, m = sl.m;
  const long long totF = sl.F * sl.n_groups, totU = sl.U * sl.n_groups;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < totF + totU;
       e += (long long)gridDim.x * blockDim.x) {
    if (e < totF) {
      const long long sg = e / sl.F, o = e % sl.F;
      const long long g = group_map ? group_map[sg] : sg;
      if (g < 0) continue;
      const double v = sf[e];
      if (v == 0.0) continue;
      long long d;
      if (o < sl.numcat_base) {
        d = o;
      } else {
        const long long q = o - sl.numcat_base, i = q / sl.total_dom, t = q % sl.total_dom;
        int c = 0;
        while (c + 1 < m && t >= sl.cat_off[c + 1]) c++;
        const long long ss = t - sl.cat_off[c];
        const long long slot = tr.col[c] ? (long long)tr.col[c][ss] : ss + (sl.lo[c] - dl.lo[c]);
        if (slot < 0) continue;
        d = dl.numcat_base + i * dl.total_dom + dl.cat_off[c] + slot;
      }
      df[g * dl.F + d] += v;
    } else {
      const long long e2 = e - totF, sg = e2 / sl.U, o = e2 % sl.U;
      const long long g = group_map ? group_map[sg] : sg;
      if (g < 0) continue;
      const unsigned long long v = su[e2];
      if (v == 0ull) continue;
      long long d;
      if (o == 0) {
        d = 0;
      } else if (o < sl.pair_base) {
        const long long t = o - 1;
        int c = 0;
        while (c + 1 < m && t >= sl.cat_off[c + 1]) c++;
        const long long ss = t - sl.cat_off[c];
        const long long slot = tr.col[c] ? (long long)tr.col[c][ss] : ss + (sl.lo[c] - dl.lo[c]);
        if (slot < 0) continue;
        d = 1 + dl.cat_off[c] + slot;
      } else {
        const long long q = o - sl.pair_base;
        int k = 0, l = 1;
        // find the pair table that contains q (tables are laid out in (k<l) row-major order)
        for (int a = 0; a < m; a++)
          for (int b = a + 1; b < m; b++)
            if (q >= sl.pair_off[a * m + b]) {
              k = a;
              l = b;
            }
        const long long w = q - sl.pair_off[k * m + l];
        const long long wk = w / sl.dom[l], wl = w % sl.dom[l];
        const long long sk = tr.col[k] ? (long long)tr.col[k][wk] : wk + (sl.lo[k] - dl.lo[k]);
        const long long s2 = tr.col[l] ? (long long)tr.col[l][wl] : wl + (sl.lo[l] - dl.lo[l]);
        if (sk < 0 || s2 < 0) continue;
        if (dl.pairs_hashed) {
          if (!pair_hash_add(dhash, g, pair_key(k * m + l, sk, s2), v)) atomicExch(err, 3);
          continue;
        }
        d = dl.pair_base + dl.pair_off[k * m + l] + sk * dl.dom[l] + s2;
      }
      du[g * dl.U + d] += v;
    }
  }
  (void)n;
}

// ---------------------------------------------------------------------------------------
// Pair counts of a HASHED source state added into dst (dense or hashed), with the slot shift of
// a domain change: combine, domain growth and hash-table growth all go through here.
__global__ void __launch_bounds__(256)
    pair_hash_drain_kernel(const PairHash src, const Layout *__restrict__ dl_g, const Layout *__restrict__ sl_g,
                           unsigned long long *__restrict__ du, const PairHash dhash, int *__restrict__ err,
                           const SlotTrans tr, const int *__restrict__ group_map, int src_groups) {
  __shared__ int s_dlo[kMaxCat], s_slo[kMaxCat], s_ddom[kMaxCat];
  __shared__ long long s_poff[kMaxCat * kMaxCat];
  __shared__ int s_m, s_G, s_dhashed;
  __shared__ long long s_U, s_pbase;
  if (threadIdx.x == 0) {
    s_m = dl_g->m;
    s_G = src_groups;
    s_dhashed = dl_g->pairs_hashed;
    s_U = dl_g->U;
    s_pbase = dl_g->pair_base;
  }
  for (int i = threadIdx.x; i < kMaxCat; i += blockDim.x) {
    s_dlo[i] = dl_g->lo[i];
    s_slo[i] = sl_g->lo[i];
    s_ddom[i] = dl_g->dom[i];
  }
  for (int i = threadIdx.x; i < kMaxCat * kMaxCat; i += blockDim.x) s_poff[i] = dl_g->pair_off[i];
  __syncthreads();
  const unsigned long long total = src.capacity * (unsigned long long)s_G;
  const unsigned long long slot_mask = (1ull << kPairSlotBits) - 1;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long key = src.keys[i];
    if (key == kPairEmpty) continue;
    const unsigned long long v = src.counts[i];
    if (!v) continue;
    const long long sg = (long long)(i / src.capacity);
    const long long g = group_map ? group_map[sg] : sg;
    if (g < 0) continue;
    const int p = (int)(key >> (2 * kPairSlotBits));
    const int k = p / s_m, l = p % s_m;
    const long long wk = (long long)((key >> kPairSlotBits) & slot_mask), wl = (long long)(key & slot_mask);
    const long long sk = tr.col[k] ? (long long)tr.col[k][wk] : wk + (s_slo[k] - s_dlo[k]);
    const long long sl2 = tr.col[l] ? (long long)tr.col[l][wl] : wl + (s_slo[l] - s_dlo[l]);
    if (sk < 0 || sl2 < 0) continue;
    if (s_dhashed) {
      if (!pair_hash_add(dhash, g, pair_key(p, sk, sl2), v)) atomicExch(err, 3);
    } else {
      atomicAdd(du + g * s_U + s_pbase + s_poff[k * s_m + l] + sk * s_ddom[l] + sl2, v);
    }
  }
}

__global__ void pair_hash_clear_kernel(PairHash h, unsigned long long total) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    h.keys[i] = kPairEmpty;
    h.counts[i] = 0ull;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *h.n_entries = 0ull;
}

// ---------------------------------------------------------------------------------------
// sum_triple / sum_nb_agg (sum.cpp:86-149): column sums of `rows` lifted triples laid out
// row-major [rows][width] (width = n for lin, nq for quad), added into state[0..width).
// One block per group of columns; fp64 accumulation.
__global__ void __launch_bounds__(256)
    lifted_colsum_kernel(const float *__restrict__ x, unsigned long long rows, int width, double *__restrict__ state) {
  for (int col = blockIdx.x; col < width; col += gridDim.x) {
    double acc = 0.0;
    for (unsigned long long r = threadIdx.x; r < rows; r += blockDim.x) acc += (double)x[r * width + col];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ double part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; w++) t += part[w];
      state[col] += t;
    }
    __syncthreads();
  }
}

// N += sum of the lifted rows' N (sum.cpp:86-89)
__global__ void __launch_bounds__(256) lifted_count_kernel(const int32_t *__restrict__ n, unsigned long long rows,
                                                           unsigned long long *__restrict__ u64) {
  long long acc = 0;
  for (unsigned long long r = threadIdx.x; r < rows; r += blockDim.x) acc += n[r];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(u64, (unsigned long long)acc);
}

// Sparse entries of lifted triples (sum.cpp:197-260): tag selects the table
//   tag = c                      lin_cat   of categorical column c:        counts[key1] += value
//   tag = 64 + l*32 + c          num_cat   of numeric l, categorical c:    numcat[l][key1] += value
//   tag = 2048 + k*32 + l (k<l)  cat_cat   of the pair (k,l):              pairs[key1][key2] += value
struct LiftedEntry {
  int32_t tag, key1, key2;
  float value;
};
__global__ void __launch_bounds__(256)
    lifted_scatter_kernel(const LiftedEntry *__restrict__ e, unsigned long long n_entries, const Layout *__restrict__ lay_g,
                          double *__restrict__ f64, unsigned long long *__restrict__ u64, int *__restrict__ err,
                          const PairHash hash, int group) {
  // f64 / u64 point at the state of GROUP BY slot `group` (the hash table is partitioned by group)
  __shared__ Layout lay;
  {
    const int *src = reinterpret_cast<const int *>(lay_g);
    int *dst = reinterpret_cast<int *>(&lay);
    for (int i = threadIdx.x; i < (int)(sizeof(Layout) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const int m = lay.m;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_entries;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const LiftedEntry x = e[i];
    if (x.tag < 64) {
      const int c = x.tag;
      const long long s = (long long)x.key1 - lay.lo[c];
      if (c >= m || s < 0 || s >= lay.dom[c]) {
        atomicExch(err, 1);
        continue;
      }
      atomicAdd(u64 + 1 + lay.cat_off[c] + s, (unsigned long long)llrintf(x.value));
    } else if (x.tag < 2048) {
      const int l = (x.tag - 64) / 32, c = (x.tag - 64) % 32;
      const long long s = (long long)x.key1 - lay.lo[c];
      if (c >= m || l >= lay.n || s < 0 || s >= lay.dom[c]) {
        atomicExch(err, 1);
        continue;
      }
      atomicAdd(f64 + lay.numcat_base + (long long)l * lay.total_dom + lay.cat_off[c] + s, (double)x.value);
    } else {
      const int k = (x.tag - 2048) / 32, l = (x.tag - 2048) % 32;
      const long long sk = (long long)x.key1 - lay.lo[k], sl = (long long)x.key2 - lay.lo[l];
      if (k >= l || l >= m || sk < 0 || sk >= lay.dom[k] || sl < 0 || sl >= lay.dom[l]) {
        atomicExch(err, 1);
        continue;
      }
      if (!lay.pairs_hashed)
        atomicAdd(u64 + lay.pair_base + lay.pair_off[k * m + l] + sk * lay.dom[l] + sl, (unsigned long long)llrintf(x.value));
      else if (!pair_hash_add(hash, group, pair_key(k * m + l, sk, sl), (unsigned long long)llrintf(x.value)))
        atomicExch(err, 3);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Synthetic columns for tests and bench (counter-based, so the host can regenerate any
// slice bit-for-bit: see duckdb_imputation_b200/synth.py).
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__global__ void gen_uniform_kernel(float *out, unsigned long long n, unsigned long long seed, unsigned long long first) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint64_t h = mix64(seed * 0xD1342543DE82EF95ull + first + i);
    out[i] = (float)(h >> 40) * (1.0f / 16777216.0f);
  }
}

__global__ void gen_int_kernel(int32_t *out, unsigned long long n, unsigned long long seed, unsigned long long first,
                               int lo, unsigned int range) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint64_t h = mix64(seed * 0xD1342543DE82EF95ull + first + i);
    out[i] = lo + (int)((h >> 33) % range);
  }
}

}  // namespace cfb
