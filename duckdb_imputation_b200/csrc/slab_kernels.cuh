// slab_kernels.cuh -- kernel family K3 (categorical aggregates) and the GROUP BY path, v2.
//
// Replaces the per-row std::map updates of Triple::SumNoLift (sum_no_lift.cpp:158-214) /
// Triple::sum_to_nb_agg (sum_to_nb_agg.cpp:124-145) and, for GROUP BY scans, the per-row
// states[sdata.sel->get_index(j)] routing of every numeric update (sum_no_lift.cpp:83-147).
//
// The work is scatter-bound (C3: 155 table updates per 80-byte row), so the kernel minimises
// the NUMBER of atomic operations rather than bytes:
//   * the payload a row adds to one table row -- [count, x_0..x_{n-1}] for (column, key), and
//     [1, x, x x^T] for its GROUP BY slot -- goes out as 128-bit vector reductions
//     (red.global.add.v4.f32, SASS REDG.E.ADD.F32x4): 4 fp32 adds per L2 atomic op;
//   * those fp32 tables are per-CTA "slabs" in global memory (L2-resident, no cross-CTA
//     contention) that are folded into the fp64 / uint64 context state every `flush_tiles`
//     tiles, which also bounds the length of every fp32 run;
//   * (key1,key2) pair counts are one 64-bit integer reduction each, straight into the state.
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

#include "pair_hash.cuh"
#include "state_layout.h"

namespace cfb {

constexpr int kSlabThreads = 256;
constexpr int kSlabRowsPerThread = 4;
constexpr int kSlabTile = kSlabThreads * kSlabRowsPerThread;

__device__ __forceinline__ void red_v4(float *p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_f32(float *p, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}
__device__ __forceinline__ void red_u64(unsigned long long *p, unsigned long long a) {
  asm volatile("red.global.add.u64 [%0], %1;" ::"l"(p), "l"(a) : "memory");
}

__host__ __device__ constexpr int pad4(int v) { return (v + 3) & ~3; }

// Per-CTA slab: [ G x VN numeric floats | G x total_dom x P categorical floats ]
//   VN = pad4(1 + n + nq) when the scan is grouped (do_numeric), else 0
//   P  = pad4(1 + n) for the triple ring, 1 for the NB ring (count only)
struct SlabShape {
  int vn, p;
  long long floats;  // per CTA
};
__host__ __device__ inline SlabShape slab_shape(const Layout &L, int do_numeric) {
  SlabShape s;
  s.vn = do_numeric ? pad4(1 + L.n + L.nq) : 0;
  s.p = L.kind == 0 ? pad4(1 + L.n) : 1;
  s.floats = (long long)L.n_groups * s.vn + (long long)L.n_groups * L.total_dom * s.p;
  s.floats = (s.floats + 3) & ~3ll;
  return s;
}

struct SlabArgs {
  ScanCols cols;
  const Layout *lay;
  unsigned long long n_rows;
  int do_numeric;   // GROUP BY scan: N / lin / quad go through the slab too
  int flush_tiles;  // fold the slab into the state every this many tiles
  float *slab;      // [gridDim.x][shape.floats], all zero on entry and on exit
  double *f64;
  unsigned long long *u64;
  int *err;
  PairHash hash;    // used when lay->pairs_hashed
};

template <int B, int E, class F>
__device__ __forceinline__ void slab_for(F &&f) {
  if constexpr (B < E) {
    f(std::integral_constant<int, B>{});
    slab_for<B + 1, E>(f);
  }
}
__host__ __device__ constexpr int tri_row(int n, int p) {
  int i = 0;
  while (p >= n - i) {
    p -= n - i;
    i++;
  }
  return i;
}
__host__ __device__ constexpr int tri_col(int n, int p) {
  int i = 0;
  while (p >= n - i) {
    p -= n - i;
    i++;
  }
  return i + p;
}
// element T of the numeric payload [1 | x_0..x_{N-1} | quad | 0 padding]; pay = [1, x...]
template <int N, int KIND, int T>
__device__ __forceinline__ float numeric_value(const float *pay) {
  constexpr int NQ = KIND == 0 ? N * (N + 1) / 2 : N;
  if constexpr (T <= N) {
    return pay[T];
  } else if constexpr (T <= N + NQ) {
    if constexpr (KIND == 0) {
      constexpr int i = tri_row(N, T - 1 - N), j = tri_col(N, T - 1 - N);
      return pay[1 + i] * pay[1 + j];
    } else {
      return pay[1 + (T - 1 - N)] * pay[1 + (T - 1 - N)];
    }
  } else {
    return 0.f;
  }
}

template <int N, int KIND>
__global__ void __launch_bounds__(kSlabThreads) slab_scan_kernel(const __grid_constant__ SlabArgs a) {
  __shared__ Layout lay;
  __shared__ int s_slot[kMaxCat][kSlabThreads];
  {
    const int *src = reinterpret_cast<const int *>(a.lay);
    int *dst = reinterpret_cast<int *>(&lay);
    for (int i = threadIdx.x; i < (int)(sizeof(Layout) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  constexpr int NQ = KIND == 0 ? N * (N + 1) / 2 : N;
  constexpr int P = KIND == 0 ? pad4(1 + N) : 1;
  const int m = lay.m;
  const int vn = a.do_numeric ? pad4(1 + N + NQ) : 0;
  const long long cat_base = (long long)lay.n_groups * vn;
  long long slab_floats = cat_base + (long long)lay.n_groups * lay.total_dom * P;
  slab_floats = (slab_floats + 3) & ~3ll;
  float *slab = a.slab + (size_t)blockIdx.x * slab_floats;

  // fold this CTA's fp32 slab into the fp64 / u64 state and zero it again
  auto flush = [&]() {
    __threadfence();
    __syncthreads();
    for (long long i = threadIdx.x; i < slab_floats; i += blockDim.x) {
      const float v = __ldcg(slab + i);
      if (v == 0.f) continue;
      __stcg(slab + i, 0.f);
      if (i < cat_base) {
        const long long g = i / vn;
        const int j = (int)(i % vn);
        if (j == 0)
          red_u64(a.u64 + g * lay.U, (unsigned long long)v);
        else if (j <= N + NQ)
          atomicAdd(a.f64 + g * lay.F + (j - 1), (double)v);
      } else {
        const long long q = i - cat_base, per_g = lay.total_dom * P;
        const long long g = q / per_g, rem = q % per_g, t = rem / P;
        const int j = (int)(rem % P);
        if (j == 0)
          red_u64(a.u64 + g * lay.U + 1 + t, (unsigned long long)v);
        else if (j <= N)
          atomicAdd(a.f64 + g * lay.F + lay.numcat_base + (long long)(j - 1) * lay.total_dom + t, (double)v);
      }
    }
    __threadfence();
    __syncthreads();
  };

  const unsigned long long n_tiles = (a.n_rows + kSlabTile - 1) / kSlabTile;
  int since_flush = 0;
  for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#pragma unroll 1
    for (int rr = 0; rr < kSlabRowsPerThread; rr++) {
      const unsigned long long r = tile * kSlabTile + (unsigned long long)rr * kSlabThreads + threadIdx.x;
      if (r >= a.n_rows) break;
      int g = 0;
      if (a.cols.group) {
        g = a.cols.group[r];
        if (g < 0) continue;  // filtered row (e.g. WHERE col_IS_NULL IS FALSE)
        if (g >= lay.n_groups) {
          atomicExch(a.err, 2);
          continue;
        }
      }
      // payload [1, x_0 .. x_{N-1}] padded to a multiple of 4
      float pay[pad4(1 + N)];
      pay[0] = 1.f;
#pragma unroll
      for (int k = 0; k < N; k++) pay[1 + k] = a.cols.num[k][r];
#pragma unroll
      for (int k = 1 + N; k < pad4(1 + N); k++) pay[k] = 0.f;

      if (a.do_numeric) {
        // [1 | lin | quad] of this row into its group's slab row, 4 values per reduction
        float *dst = slab + (long long)g * vn;
        slab_for<0, pad4(1 + N + NQ) / 4>([&](auto q_) {
          constexpr int q = decltype(q_)::value;
          red_v4(dst + 4 * q, numeric_value<N, KIND, 4 * q>(pay), numeric_value<N, KIND, 4 * q + 1>(pay),
                 numeric_value<N, KIND, 4 * q + 2>(pay), numeric_value<N, KIND, 4 * q + 3>(pay));
        });
      }
      if (m == 0) continue;
      bool ok = true;
      for (int c = 0; c < m; c++) {
        const long long d = (long long)a.cols.cat[c][r] - (long long)lay.lo[c];
        if (d < 0 || d >= lay.dom[c]) ok = false;
        s_slot[c][threadIdx.x] = (int)d;
      }
      if (!ok) {
        atomicExch(a.err, 1);
        continue;
      }
      float *cat_slab = slab + cat_base + (long long)g * lay.total_dom * P;
      for (int c = 0; c < m; c++) {
        float *dst = cat_slab + (lay.cat_off[c] + s_slot[c][threadIdx.x]) * P;
        if constexpr (KIND == 0) {
#pragma unroll
          for (int k = 0; k < P; k += 4) red_v4(dst + k, pay[k], pay[k + 1], pay[k + 2], pay[k + 3]);
        } else {
          red_f32(dst, 1.f);
        }
      }
      if constexpr (KIND == 0) {
        if (!lay.pairs_hashed) {
          unsigned long long *pairs = a.u64 + (long long)g * lay.U + lay.pair_base;
          for (int k = 0; k < m; k++) {
            const long long sk = s_slot[k][threadIdx.x];
            for (int l = k + 1; l < m; l++)
              red_u64(pairs + lay.pair_off[k * m + l] + sk * lay.dom[l] + s_slot[l][threadIdx.x], 1ull);
          }
        } else {
          for (int k = 0; k < m; k++) {
            const long long sk = s_slot[k][threadIdx.x];
            for (int l = k + 1; l < m; l++)
              if (!pair_hash_add(a.hash, g, pair_key(k * m + l, sk, s_slot[l][threadIdx.x]), 1ull)) atomicExch(a.err, 3);
          }
        }
      }
    }
    if (++since_flush >= a.flush_tiles) {
      flush();
      since_flush = 0;
    }
  }
  flush();
}

}  // namespace cfb
