// gram_kernel.cuh -- kernel families K1+K2 of the cofactor aggregate for sm_100a:
//   count / lin_agg column sums and the quad_agg Gram matrix X^T X (packed upper triangle),
//   or, for the Naive-Bayes ring, lin_agg and the diagonal sum x_k^2 only.
//
// Replaces the numeric half of Triple::SumNoLift (sum_no_lift.cpp:119-147) and of
// Triple::sum_to_nb_agg (sum_to_nb_agg.cpp:98-117) for one ungrouped scan of device-resident
// columnar (SoA) input.
//
// Shape of the kernel (one persistent CTA per SM):
//   * warp P (producer): one elected lane streams row tiles of every FLOAT column into a
//     kStages-deep shared-memory ring with 1-D bulk async copies (TMA engine, SASS UBLKCP),
//     completion counted on per-stage "full" mbarriers, L2 evict-first (data is read once);
//   * consumer warps: wait on "full", accumulate with packed fp32x2 FMAs (SASS FFMA2: two rows
//     per lane per instruction), release the stage through the "empty" mbarrier.
//     The n(n+1)/2 pair accumulators do not fit one thread, so the upper triangle is cut into
//     kRoles rectangles ("roles"); the warps of one role group share the same rows and each
//     owns one rectangle.  Role groups split the rows of a tile.
//   * fp32 partials only live for a bounded run of rows (kFlushTiles tiles); they are then
//     warp-reduced and added into fp64 shared-memory totals.  At the end each CTA writes its
//     fp64 partial vector, and the last CTA to finish adds all partials in CTA order into the
//     context state -- deterministic, no floating-point atomics.
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

#include "ptx_sm100.cuh"

namespace cfb {

constexpr int kMaxCols = 32;

struct NumCols {
  const float *p[kMaxCols];
};

__device__ __forceinline__ void ffma2(unsigned long long &d, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ void fadd2(unsigned long long &d, unsigned long long a) {
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(d) : "l"(a));
}
__device__ __forceinline__ unsigned long long lds64(const float *p) {
  return *reinterpret_cast<const unsigned long long *>(p);
}
struct __align__(16) U64x2 {
  unsigned long long lo, hi;  // rows (r, r+1) and (r+2, r+3) of one column
};
__device__ __forceinline__ U64x2 lds128(const float *p) { return *reinterpret_cast<const U64x2 *>(p); }
__device__ __forceinline__ float hsum2(unsigned long long v) {
  return __uint_as_float((unsigned)(v & 0xffffffffull)) + __uint_as_float((unsigned)(v >> 32));
}

// ------------------------------------------------------------------ role geometry
// A role owns the pairs (i, j), i0 <= i < i1, j0 <= j < j1, i <= j, plus the column sums
// (lin_agg) of the columns in [la0, la1) and [lb0, lb1) -- all of which it loads anyway.
struct RoleGeom {
  int i0, i1, j0, j1, la0, la1, lb0, lb1;
};

template <int N, bool DIAG>
struct GramShape {
  static constexpr int kRoles = DIAG ? 1 : (N <= 10 ? 1 : (N <= 20 ? 4 : 10));
  // one role: 8 row groups (4 for the wide NB shapes, whose 1024-row stages would not fit twice)
  static constexpr int kGroups = kRoles == 1 ? (N <= 16 ? 8 : 4) : (kRoles == 4 ? 2 : 1);
  static constexpr int kConsumerWarps = kRoles * kGroups;
  // 8 consumer warps = 2 warpgroups; the producer warp sits in a third warpgroup (3 idle warps)
  // so that setmaxnreg can move its registers to the consumers.
  static constexpr bool kRegSplit = kConsumerWarps == 8;
  static constexpr int kThreads = kRegSplit ? 384 : (kConsumerWarps + 1) * 32;
  static constexpr int kOut = DIAG ? 2 * N : N + N * (N + 1) / 2;  // [lin | quad]

  static constexpr RoleGeom role(int r) {
    if (kRoles == 1) return RoleGeom{0, N, 0, N, 0, N, 0, 0};
    if (kRoles == 4) {
      const int h = (N + 1) / 2, q = (h + 1) / 2, hm = h + (N - h) / 2;
      if (r == 0) return RoleGeom{0, h, 0, h, 0, 0, 0, 0};
      if (r == 1) return RoleGeom{h, N, h, N, 0, 0, 0, 0};
      if (r == 2) return RoleGeom{0, q, h, N, 0, q, h, hm};
      return RoleGeom{q, h, h, N, q, h, hm, N};
    }
    // 10 roles: 4 column groups, blocks (gi <= gj) in row-major order
    int gi = 0, gj = 0, k = 0;
    for (int a = 0; a < 4; a++)
      for (int b = a; b < 4; b++, k++)
        if (k == r) {
          gi = a;
          gj = b;
        }
    const int bi0 = gi * N / 4, bi1 = (gi + 1) * N / 4, bj0 = gj * N / 4, bj1 = (gj + 1) * N / 4;
    if (gi == gj) return RoleGeom{bi0, bi1, bj0, bj1, bi0, bi1, 0, 0};
    return RoleGeom{bi0, bi1, bj0, bj1, 0, 0, 0, 0};
  }
};

// packed upper-triangle index of (i <= j): i*n - i*(i+1)/2 + j   (ML/utils.cpp:195-197)
__host__ __device__ constexpr int tri_index(int n, int i, int j) { return i * n - i * (i + 1) / 2 + j; }

template <int N, bool DIAG, int R>
struct Role {
  using S = GramShape<N, DIAG>;
  static constexpr int i0 = S::role(R).i0, i1 = S::role(R).i1, j0 = S::role(R).j0, j1 = S::role(R).j1;
  static constexpr int la0 = S::role(R).la0, la1 = S::role(R).la1, lb0 = S::role(R).lb0, lb1 = S::role(R).lb1;
  static constexpr int kI = i1 - i0;
  static constexpr int kJ = j1 - j0;
  static constexpr bool kSame = (i0 == j0 && i1 == j1);
  static constexpr int kLinA = la1 - la0, kLinB = lb1 - lb0;
  // does this role own pair (a, b) (role-local coordinates)?
  static constexpr bool owns(int a, int b) { return DIAG ? (kSame && a == b) : (j0 + b >= i0 + a); }
  // accumulator slot of pair (a, b): its rank among the owned pairs in row-major order
  static constexpr int slot(int a, int b) {
    int c = 0;
    for (int i = 0; i < kI; i++)
      for (int j = 0; j < kJ; j++)
        if (owns(i, j)) {
          if (i == a && j == b) return c;
          c++;
        }
    return c;
  }
  static constexpr int kPairs = slot(kI, kJ);  // (kI, kJ) is never owned: counts all pairs
  static constexpr int kAcc = kPairs + kLinA + kLinB;
  // canonical output index of pair (a, b): [lin | quad]
  static constexpr int out_index(int a, int b) {
    return DIAG ? (N + i0 + a) : (N + tri_index(N, i0 + a, j0 + b));
  }
};

template <int B, int E, class F>
__device__ __forceinline__ void static_for(F &&f) {
  if constexpr (B < E) {
    f(std::integral_constant<int, B>{});
    static_for<B + 1, E>(f);
  }
}

struct GramArgs {
  NumCols cols;
  unsigned long long n_rows;  // all rows of the scan
  int stages;                 // depth of the shared-memory ring
  int flush_tiles;            // fp32 partials are folded into fp64 every this many tiles
  double *partials;           // [gridDim.x][kOut]
  double *state;              // [kOut] context totals (lin | quad), += at the end
  unsigned int *ticket;       // zero-initialised; reset by the last CTA
  unsigned long long *count;  // N of the context: += n_rows by the last CTA (nullptr: not wanted)
};

// Rows per ring stage.  A tile is consumed in 128-row warp iterations split over kGroups row
// groups, so kRows is a multiple of 128 * kGroups (no idle warps); narrow tables get longer tiles so
// that a column chunk (one bulk copy) stays >= 2 KB; 768 rows measured best for the 4-role shape.
#ifndef CFB_TR_WIDE
#define CFB_TR_WIDE 768
#endif
template <int N, bool DIAG>
struct GramTile {
  static constexpr int kGroups = GramShape<N, DIAG>::kGroups;
  static constexpr int kRows = kGroups == 8 ? (N <= 4 ? 2048 : 1024) : (kGroups == 4 ? 512 : (kGroups == 2 ? CFB_TR_WIDE : 512));
  static_assert(kRows % (128 * kGroups) == 0, "tile must split evenly over the row groups");
};

template <int N, bool DIAG>
__host__ __device__ constexpr size_t gram_smem_bytes(int stages) {
  return (size_t)stages * N * GramTile<N, DIAG>::kRows * sizeof(float) +
         (size_t)GramShape<N, DIAG>::kGroups * GramShape<N, DIAG>::kOut * sizeof(double) +
         2 * (size_t)stages * sizeof(uint64_t);
}

__device__ __forceinline__ float warp_sum(float t) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}

template <int N, bool DIAG, int R, int TR>
__device__ __forceinline__ void gram_consume(const GramArgs &a, const float *ring, double *totals,
                                             uint64_t *full, uint64_t *empty, int group, int lane) {
  using S = GramShape<N, DIAG>;
  using RL = Role<N, DIAG, R>;
  unsigned long long acc[RL::kAcc];
#pragma unroll
  for (int i = 0; i < RL::kAcc; i++) acc[i] = 0ull;

  const unsigned long long rows4 = a.n_rows & ~3ull;  // rows delivered through the ring
  const unsigned long long n_tiles = (rows4 + TR - 1) / TR;
  int stage = 0;
  uint32_t phase = 0;
  int since_flush = 0;

  // fp32 partials -> fp64 totals: warp-reduce every accumulator, one lane adds it.
  auto flush = [&]() {
    static_for<0, RL::kI>([&](auto ai_) {
      static_for<0, RL::kJ>([&](auto bj_) {
        constexpr int ai = decltype(ai_)::value, bj = decltype(bj_)::value;
        if constexpr (RL::owns(ai, bj)) {
          constexpr int s = RL::slot(ai, bj);
          constexpr int out = RL::out_index(ai, bj);
          const float t = warp_sum(hsum2(acc[s]));
          if (lane == (s & 31)) totals[out] += (double)t;
          acc[s] = 0ull;
        }
      });
    });
    static_for<0, RL::kLinA>([&](auto k_) {
      constexpr int k = decltype(k_)::value;
      const float t = warp_sum(hsum2(acc[RL::kPairs + k]));
      if (lane == (k & 31)) totals[RL::la0 + k] += (double)t;
      acc[RL::kPairs + k] = 0ull;
    });
    static_for<0, RL::kLinB>([&](auto k_) {
      constexpr int k = decltype(k_)::value;
      const float t = warp_sum(hsum2(acc[RL::kPairs + RL::kLinA + k]));
      if (lane == (k & 31)) totals[RL::lb0 + k] += (double)t;
      acc[RL::kPairs + RL::kLinA + k] = 0ull;
    });
    __syncwarp();
  };

  for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const unsigned long long row0 = tile * TR;
    const int valid = (int)((rows4 - row0) < (unsigned long long)TR ? (rows4 - row0) : (unsigned long long)TR);
    const float *st = ring + (size_t)stage * (N * TR);
    ptx::mbar_wait(&full[stage], phase);
#pragma unroll 1
    for (int it = group; it < TR / 128; it += S::kGroups) {
      const int row = it * 128 + lane * 4;
      if (row >= valid) break;  // valid is a multiple of 4: a lane's four rows are all in or all out
      U64x2 xi[RL::kI];
      U64x2 xj[RL::kSame ? 1 : RL::kJ];
#pragma unroll
      for (int k = 0; k < RL::kI; k++) xi[k] = lds128(st + (RL::i0 + k) * TR + row);
      if constexpr (!RL::kSame) {
#pragma unroll
        for (int k = 0; k < RL::kJ; k++) xj[k] = lds128(st + (RL::j0 + k) * TR + row);
      }
      static_for<0, RL::kI>([&](auto ai_) {
        static_for<0, RL::kJ>([&](auto bj_) {
          constexpr int ai = decltype(ai_)::value, bj = decltype(bj_)::value;
          if constexpr (RL::owns(ai, bj)) {
            constexpr int s = RL::slot(ai, bj);
            if constexpr (RL::kSame) {
              ffma2(acc[s], xi[ai].lo, xi[bj].lo);
              ffma2(acc[s], xi[ai].hi, xi[bj].hi);
            } else {
              ffma2(acc[s], xi[ai].lo, xj[bj].lo);
              ffma2(acc[s], xi[ai].hi, xj[bj].hi);
            }
          }
        });
      });
      // column sums (lin_agg) this role owns: ranges are subsets of its loaded columns
      static_for<0, RL::kLinA>([&](auto k_) {
        constexpr int k = decltype(k_)::value;
        fadd2(acc[RL::kPairs + k], xi[RL::la0 - RL::i0 + k].lo);
        fadd2(acc[RL::kPairs + k], xi[RL::la0 - RL::i0 + k].hi);
      });
      static_for<0, RL::kLinB>([&](auto k_) {
        constexpr int k = decltype(k_)::value;
        if constexpr (RL::kSame) {
          fadd2(acc[RL::kPairs + RL::kLinA + k], xi[RL::lb0 - RL::i0 + k].lo);
          fadd2(acc[RL::kPairs + RL::kLinA + k], xi[RL::lb0 - RL::i0 + k].hi);
        } else {
          fadd2(acc[RL::kPairs + RL::kLinA + k], xj[RL::lb0 - RL::j0 + k].lo);
          fadd2(acc[RL::kPairs + RL::kLinA + k], xj[RL::lb0 - RL::j0 + k].hi);
        }
      });
    }
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&empty[stage]);
    if (++stage == a.stages) {
      stage = 0;
      phase ^= 1u;
    }
    if (++since_flush == a.flush_tiles) {
      flush();
      since_flush = 0;
    }
  }
  flush();
}

template <int N, bool DIAG, int TR>
__global__ void __launch_bounds__(GramShape<N, DIAG>::kThreads, 1)
    gram_scan_kernel(const __grid_constant__ GramArgs a) {
  using S = GramShape<N, DIAG>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *ring = reinterpret_cast<float *>(smem_raw);
  double *totals = reinterpret_cast<double *>(ring + (size_t)a.stages * N * TR);  // [kGroups][kOut]
  uint64_t *full = reinterpret_cast<uint64_t *>(totals + S::kGroups * S::kOut);
  uint64_t *empty = full + a.stages;
  __shared__ unsigned int s_is_last;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < S::kGroups * S::kOut; i += blockDim.x) totals[i] = 0.0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; s++) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], S::kConsumerWarps);
    }
    ptx::fence_mbar_init();
  }
  __syncthreads();

  const unsigned long long rows4 = a.n_rows & ~3ull;
  const unsigned long long n_tiles = (rows4 + TR - 1) / TR;

  // Register split (kRegSplit): 384 threads start with 168 registers each; the producer
  // warpgroup shrinks to 24 and the two consumer warpgroups grow to 240
  // (24*128 + 240*256 = 64512 <= 65536).  Each setmaxnreg sits inside its own branch so that
  // the register budget of the code that follows is unambiguous to ptxas.
  if (warp > S::kConsumerWarps) {
    // idle warps of the producer warpgroup
    if constexpr (S::kRegSplit) asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    return;
  } else if (warp == S::kConsumerWarps) {
    if constexpr (S::kRegSplit) asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    // ------------------------------------------------------------ producer warp
    // Lane 0 waits for the stage to drain and arms its "full" barrier; then lane c issues the
    // bulk copy of column c, so the N copies of a tile are issued in parallel (one thread can
    // only issue a bulk copy every ~100 cycles: profiles/r01_stream_ceiling.txt).
    int stage = 0;
    uint32_t phase = 0;
    for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const unsigned long long row0 = tile * TR;
      const uint32_t valid = (uint32_t)((rows4 - row0) < (unsigned long long)TR ? (rows4 - row0) : (unsigned long long)TR);
      if (lane == 0) {
        ptx::mbar_wait(&empty[stage], phase ^ 1u);
        ptx::mbar_arrive_expect_tx(&full[stage], valid * (uint32_t)sizeof(float) * N);
      }
      __syncwarp();
      float *dst = ring + (size_t)stage * (N * TR);
      if (lane < N) ptx::bulk_g2s_plain(dst + lane * TR, a.cols.p[lane] + row0, valid * (uint32_t)sizeof(float), &full[stage]);
      if (++stage == a.stages) {
        stage = 0;
        phase ^= 1u;
      }
    }
    return;
  } else {
    // ----------------------------------------------------------- consumer warps
    if constexpr (S::kRegSplit) asm volatile("setmaxnreg.inc.sync.aligned.u32 240;");
    const int role = warp % S::kRoles, group = warp / S::kRoles;
    double *tot = totals + group * S::kOut;
    if constexpr (S::kRoles == 1) {
      gram_consume<N, DIAG, 0, TR>(a, ring, tot, full, empty, group, lane);
    } else if constexpr (S::kRoles == 4) {
      switch (role) {
        case 0: gram_consume<N, DIAG, 0, TR>(a, ring, tot, full, empty, group, lane); break;
        case 1: gram_consume<N, DIAG, 1, TR>(a, ring, tot, full, empty, group, lane); break;
        case 2: gram_consume<N, DIAG, 2, TR>(a, ring, tot, full, empty, group, lane); break;
        default: gram_consume<N, DIAG, 3, TR>(a, ring, tot, full, empty, group, lane); break;
      }
    } else {
      switch (role) {
        case 0: gram_consume<N, DIAG, 0, TR>(a, ring, tot, full, empty, group, lane); break;
        case 1: gram_consume<N, DIAG, 1, TR>(a, ring, tot, full, empty, group, lane); break;
        case 2: gram_consume<N, DIAG, 2, TR>(a, ring, tot, full, empty, group, lane); break;
        case 3: gram_consume<N, DIAG, 3, TR>(a, ring, tot, full, empty, group, lane); break;
        case 4: gram_consume<N, DIAG, 4, TR>(a, ring, tot, full, empty, group, lane); break;
        case 5: gram_consume<N, DIAG, 5, TR>(a, ring, tot, full, empty, group, lane); break;
        case 6: gram_consume<N, DIAG, 6, TR>(a, ring, tot, full, empty, group, lane); break;
        case 7: gram_consume<N, DIAG, 7, TR>(a, ring, tot, full, empty, group, lane); break;
        case 8: gram_consume<N, DIAG, 8, TR>(a, ring, tot, full, empty, group, lane); break;
        default: gram_consume<N, DIAG, 9, TR>(a, ring, tot, full, empty, group, lane); break;
      }
    }
  }
  // Only the consumer warps get here (the producer warpgroup has exited).
  constexpr int kTail = S::kConsumerWarps * 32;
  ptx::named_bar_sync(1, kTail);

  // ------------------------------------------------- CTA partial -> global, last CTA folds
  double *mine = a.partials + (size_t)blockIdx.x * S::kOut;
  for (int t = threadIdx.x; t < S::kOut; t += kTail) {
    double s = 0.0;
#pragma unroll
    for (int gq = 0; gq < S::kGroups; gq++) s += totals[gq * S::kOut + t];
    mine[t] = s;
  }
  __threadfence();
  ptx::named_bar_sync(1, kTail);
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(a.ticket, 1u);
    s_is_last = (prev == gridDim.x - 1) ? 1u : 0u;
  }
  ptx::named_bar_sync(1, kTail);
  if (s_is_last) {
    __threadfence();
    for (int t = threadIdx.x; t < S::kOut; t += kTail) {
      double s = 0.0;
      for (unsigned int b = 0; b < gridDim.x; b++) s += a.partials[(size_t)b * S::kOut + t];
      // the <= 3 rows that are not a multiple of 4 never enter the ring
      int ci, cj;
      if (t < N) {
        ci = t;
        cj = -1;
      } else if (DIAG) {
        ci = cj = t - N;
      } else {
        int p = t - N, i = 0;
        while (p >= N - i) {
          p -= N - i;
          i++;
        }
        ci = i;
        cj = i + p;
      }
      for (unsigned long long r = rows4; r < a.n_rows; r++) {
        const double xi = (double)a.cols.p[ci][r];
        s += (cj < 0) ? xi : xi * (double)a.cols.p[cj][r];
      }
      a.state[t] += s;
    }
    if (threadIdx.x == 0) {
      *a.ticket = 0u;
      if (a.count) *a.count += a.n_rows;  // (a launch of its own for one add cost C1 a tenth of its time)
    }
  }
}

}  // namespace cfb
