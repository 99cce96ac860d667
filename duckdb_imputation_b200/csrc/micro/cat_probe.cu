// cat_probe.cu -- microbenchmark for the categorical path (kernel family K3) on the C3 shape
// (10 FLOAT + 10 INT columns, domain 100): which scatter primitive is fastest on sm_100a?
//
//   P0  pair counts, L2 reductions (red.global.add.u64 per (row, pair))            -- what slab_scan does
//   P1  pair counts, shared-memory atomics (u32 tables, ATOMS), 5 tables per CTA
//   P2  pair counts, NON-atomic u16 tables in shared memory: one warp owns one table, lanes = 32 rows,
//       duplicates inside the warp are merged with match.any and the group leader does a plain
//       LDS / IADD / STS; 9 tables per CTA, 5 CTA roles cover the 45 pairs
//   S0  per-key sums [count, x_0..x_9], L2 vector reductions (red.global.add.v4.f32) -- what slab_scan does
//   S1  per-key sums, warp-private fp32 tables in shared memory, 10 rows x 3 lanes per step, each lane a
//       float4 of the payload with plain LDS.128 / FADD / STS.128; equal keys inside a step are
//       serialised by rank (match.any)
// Prints rows/s per variant; verifies P2 / S1 against P0 / S0.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int M = 10, N = 10, DOM = 100, NPAIR = 45;
constexpr int P = 12;  // payload floats per (column, key): [1, x0..x9, 0]

struct Cols {
  const int *cat[M];
  const float *num[N];
};

__device__ __forceinline__ unsigned hash32(unsigned long long x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return (unsigned)x;
}
__global__ void gen(int *cat, float *num, size_t rows, unsigned seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < rows; i += (size_t)gridDim.x * blockDim.x) {
    unsigned h = hash32(i * 1315423911ull + seed);
    if (cat) cat[i] = h % DOM;
    if (num) num[i] = (h >> 8) * (1.0f / 16777216.0f);
  }
}

// ------------------------------------------------------------------ P0: L2 reductions
__global__ void p0_kernel(Cols c, size_t rows, unsigned long long *pairs) {
  for (size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x; r < rows; r += (size_t)gridDim.x * blockDim.x) {
    int key[M];
#pragma unroll
    for (int k = 0; k < M; k++) key[k] = c.cat[k][r];
    int p = 0;
#pragma unroll
    for (int k = 0; k < M; k++)
#pragma unroll
      for (int l = k + 1; l < M; l++, p++)
        asm volatile("red.global.add.u64 [%0], %1;" ::"l"(pairs + (size_t)p * DOM * DOM + key[k] * DOM + key[l]), "l"(1ull) : "memory");
  }
}

// ------------------------------------------------------------------ P1: ATOMS u32, 5 tables per CTA, 9 roles
constexpr int P1_TABLES = 5, P1_ROLES = 9;
__global__ void __launch_bounds__(512) p1_kernel(Cols c, size_t rows, unsigned long long *pairs, int chunk_rows) {
  extern __shared__ unsigned smem_u32[];
  const int role = blockIdx.x % P1_ROLES, rep = blockIdx.x / P1_ROLES, nrep = gridDim.x / P1_ROLES;
  if (rep >= nrep) return;
  int pk[P1_TABLES], pl[P1_TABLES];
  {
    int p = 0;
    for (int k = 0; k < M; k++)
      for (int l = k + 1; l < M; l++, p++)
        if (p / P1_TABLES == role) { pk[p % P1_TABLES] = k; pl[p % P1_TABLES] = l; }
  }
  const size_t nchunks = (rows + chunk_rows - 1) / chunk_rows;
  for (int i = threadIdx.x; i < P1_TABLES * DOM * DOM; i += blockDim.x) smem_u32[i] = 0;
  __syncthreads();
  for (size_t ch = rep; ch < nchunks; ch += nrep) {
    const size_t lo = ch * chunk_rows, hi = min(rows, lo + (size_t)chunk_rows);
    for (size_t r = lo + threadIdx.x; r < hi; r += blockDim.x) {
#pragma unroll
      for (int t = 0; t < P1_TABLES; t++) atomicAdd(&smem_u32[t * DOM * DOM + c.cat[pk[t]][r] * DOM + c.cat[pl[t]][r]], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < P1_TABLES * DOM * DOM; i += blockDim.x) {
    const unsigned v = smem_u32[i];
    if (v) atomicAdd(pairs + (size_t)(role * P1_TABLES + i / (DOM * DOM)) * DOM * DOM + i % (DOM * DOM), (unsigned long long)v);
  }
}

// ------------------------------------------------------------------ P2: match.any + plain u16 RMW
constexpr int P2_TABLES = 9, P2_ROLES = 5;
template <bool USE_MATCH>
__global__ void __launch_bounds__(P2_TABLES * 32) p2_kernel(Cols c, size_t rows, unsigned long long *pairs, int chunk_rows) {
  extern __shared__ unsigned short smem_u16[];
  const int role = blockIdx.x % P2_ROLES, rep = blockIdx.x / P2_ROLES, nrep = gridDim.x / P2_ROLES;
  if (rep >= nrep) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = role * P2_TABLES + warp;
  int k = 0, l = 0;
  {
    int p = 0;
    for (int a = 0; a < M; a++)
      for (int b = a + 1; b < M; b++, p++)
        if (p == pair) { k = a; l = b; }
  }
  unsigned short *tbl = smem_u16 + warp * DOM * DOM;
  const int *ck = c.cat[k], *cl = c.cat[l];
  const size_t nchunks = (rows + chunk_rows - 1) / chunk_rows;
  for (int i = lane; i < DOM * DOM; i += 32) tbl[i] = 0;
  __syncwarp();
  for (size_t ch = rep; ch < nchunks; ch += nrep) {
    const size_t lo = ch * chunk_rows, hi = min(rows, lo + (size_t)chunk_rows);
#pragma unroll 4
    for (size_t r0 = lo; r0 < hi; r0 += 32) {
      const size_t r = r0 + lane;
      const bool live = r < hi;
      unsigned idx = 0xffff0000u | lane;
      if (live) idx = ck[r] * DOM + cl[r];
      if (USE_MATCH) {
        const unsigned mask = __match_any_sync(0xffffffffu, idx);
        const bool leader = (__ffs(mask) - 1) == lane;
        if (live && leader) tbl[idx] = (unsigned short)(tbl[idx] + __popc(mask));
      } else {  // (wrong under duplicates: only to price the match instruction)
        if (live) tbl[idx] = (unsigned short)(tbl[idx] + 1);
      }
    }
    __syncwarp();
    // fold this warp's table into the global pair counts (chunk_rows <= 65535 bounds every cell)
    unsigned long long *dst = pairs + (size_t)pair * DOM * DOM;
    for (int i = lane; i < DOM * DOM; i += 32) {
      const unsigned v = tbl[i];
      if (v) {
        asm volatile("red.global.add.u64 [%0], %1;" ::"l"(dst + i), "l"((unsigned long long)v) : "memory");
        tbl[i] = 0;
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ P3: ATOMS on packed u16 pairs, 9 tables per CTA, 5 roles
template <int THREADS>
__global__ void __launch_bounds__(THREADS) p3_kernel(Cols c, size_t rows, unsigned long long *pairs, int chunk_rows) {
  extern __shared__ unsigned smem_u32[];  // 9 tables x 5000 words (two u16 counters per word)
  const int role = blockIdx.x % P2_ROLES, rep = blockIdx.x / P2_ROLES, nrep = gridDim.x / P2_ROLES;
  if (rep >= nrep) return;
  int pk[P2_TABLES], pl[P2_TABLES];
  {
    int p = 0;
    for (int k = 0; k < M; k++)
      for (int l = k + 1; l < M; l++, p++)
        if (p / P2_TABLES == role) { pk[p % P2_TABLES] = k; pl[p % P2_TABLES] = l; }
  }
  const size_t nchunks = (rows + chunk_rows - 1) / chunk_rows;
  constexpr int WORDS = P2_TABLES * DOM * DOM / 2;
  for (int i = threadIdx.x; i < WORDS; i += THREADS) smem_u32[i] = 0;
  __syncthreads();
  for (size_t ch = rep; ch < nchunks; ch += nrep) {
    const size_t lo = ch * chunk_rows, hi = min(rows, lo + (size_t)chunk_rows);
    for (size_t r = lo + threadIdx.x; r < hi; r += THREADS) {
      int key[M];
#pragma unroll
      for (int k = 0; k < M; k++) key[k] = c.cat[k][r];
#pragma unroll
      for (int t = 0; t < P2_TABLES; t++) {
        const unsigned idx = t * DOM * DOM + key[pk[t]] * DOM + key[pl[t]];
        atomicAdd(&smem_u32[idx >> 1], 1u << ((idx & 1) * 16));
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < WORDS; i += THREADS) {
      const unsigned v = smem_u32[i];
      if (v) {
        smem_u32[i] = 0;
        unsigned long long *dst = pairs + (size_t)role * P2_TABLES * DOM * DOM + 2 * (size_t)i;
        if (v & 0xffffu) asm volatile("red.global.add.u64 [%0], %1;" ::"l"(dst), "l"((unsigned long long)(v & 0xffffu)) : "memory");
        if (v >> 16) asm volatile("red.global.add.u64 [%0], %1;" ::"l"(dst + 1), "l"((unsigned long long)(v >> 16)) : "memory");
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ S0: L2 vector reductions
__global__ void s0_kernel(Cols c, size_t rows, float *sums) {  // sums[M][DOM][P] fp32 (probe only)
  for (size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x; r < rows; r += (size_t)gridDim.x * blockDim.x) {
    float pay[P];
    pay[0] = 1.f;
#pragma unroll
    for (int j = 0; j < N; j++) pay[1 + j] = c.num[j][r];
    pay[11] = 0.f;
#pragma unroll
    for (int k = 0; k < M; k++) {
      float *dst = sums + ((size_t)k * DOM + c.cat[k][r]) * P;
#pragma unroll
      for (int q = 0; q < P; q += 4)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + q), "f"(pay[q]), "f"(pay[q + 1]), "f"(pay[q + 2]), "f"(pay[q + 3]) : "memory");
    }
  }
}

__global__ void s0b_kernel(Cols c, size_t rows, float *slabs) {  // slabs[gridDim.x][M][DOM][P]
  float *sums = slabs + (size_t)blockIdx.x * M * DOM * P;
  for (size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x; r < rows; r += (size_t)gridDim.x * blockDim.x) {
    float pay[P];
    pay[0] = 1.f;
#pragma unroll
    for (int j = 0; j < N; j++) pay[1 + j] = c.num[j][r];
    pay[11] = 0.f;
#pragma unroll
    for (int k = 0; k < M; k++) {
      float *dst = sums + ((size_t)k * DOM + c.cat[k][r]) * P;
#pragma unroll
      for (int q = 0; q < P; q += 4)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + q), "f"(pay[q]), "f"(pay[q + 1]), "f"(pay[q + 2]), "f"(pay[q + 3]) : "memory");
    }
  }
}

// ------------------------------------------------------------------ S2: CTA-shared fp32 tables, one owner warp per column
constexpr int S2_ROWS = 10;
__global__ void __launch_bounds__(M * 32) s2_kernel(Cols c, size_t rows, double *sums, int chunk_rows) {
  extern __shared__ float4 smem_f4[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;  // warp = the column it owns
  float4 *tbl = smem_f4 + (size_t)warp * DOM * 3;              // [DOM][3 float4]
  const int g = lane / 3, q = lane % 3;
  const bool lane_on = lane < 3 * S2_ROWS;
  const int *keys = c.cat[warp];
  const size_t nchunks = (rows + chunk_rows - 1) / chunk_rows;
  for (int i = lane; i < DOM * 3; i += 32) tbl[i] = make_float4(0, 0, 0, 0);
  __syncwarp();
  for (size_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const size_t lo = ch * chunk_rows, hi = min(rows, lo + (size_t)chunk_rows);
#pragma unroll 2
    for (size_t r0 = lo; r0 < hi; r0 += S2_ROWS) {
      const size_t r = r0 + g;
      const bool live = lane_on && r < hi;
      float4 pay = make_float4(0, 0, 0, 0);
      int key = 0x40000000 | lane;
      if (live) {
        key = keys[r];
        if (q == 0) pay = make_float4(1.f, c.num[0][r], c.num[1][r], c.num[2][r]);
        else if (q == 1) pay = make_float4(c.num[3][r], c.num[4][r], c.num[5][r], c.num[6][r]);
        else pay = make_float4(c.num[7][r], c.num[8][r], c.num[9][r], 0.f);
      }
      // rank among the rows of this step with the same key (no match.any: 9 shuffles over the row keys)
      unsigned rank = 0;
#pragma unroll
      for (int o = 1; o < S2_ROWS; o++) {
        const int other = __shfl_sync(0xffffffffu, key, max(lane - 3 * o, 0));
        rank += (lane >= 3 * o) && (other == key);
      }
      const unsigned maxrank = __reduce_max_sync(0xffffffffu, live ? rank : 0u);
      for (unsigned round = 0; round <= maxrank; round++) {
        if (live && rank == round) {
          float4 v = tbl[key * 3 + q];
          v.x += pay.x; v.y += pay.y; v.z += pay.z; v.w += pay.w;
          tbl[key * 3 + q] = v;
        }
        __syncwarp();
      }
    }
    for (int i = lane; i < DOM * 3; i += 32) {
      const float4 v = tbl[i];
      if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) {
        double *dst = sums + ((size_t)warp * DOM * 3 + i) * 4;
        atomicAdd(dst + 0, (double)v.x); atomicAdd(dst + 1, (double)v.y);
        atomicAdd(dst + 2, (double)v.z); atomicAdd(dst + 3, (double)v.w);
        tbl[i] = make_float4(0, 0, 0, 0);
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ S3: tile-level counting sort by key, bucket-owner threads
// No atomics on floats at all: the rows of a tile are bucketed by key per column (two ATOMS per row and
// column: count, then cursor), thread b = (column, key) then adds the payload rows of its bucket from
// shared memory into 12 registers that live across tiles.
constexpr int S3_T = 3072, S3_THREADS = 1024;
__global__ void __launch_bounds__(S3_THREADS, 1) s3_kernel(Cols c, size_t rows, double *sums, int flush_tiles) {
  extern __shared__ float4 smem_f4[];
  float4 *pay = smem_f4;                                                     // [S3_T][3]
  unsigned short *ids = reinterpret_cast<unsigned short *>(pay + S3_T * 3);  // [M][S3_T]
  unsigned *off = reinterpret_cast<unsigned *>(ids + M * S3_T);              // [M*DOM] exclusive starts
  unsigned *cur = off + M * DOM;                                             // [M*DOM] counts, then cursors
  const int tid = threadIdx.x;
  float4 acc[3] = {make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0)};
  const size_t ntiles = (rows + S3_T - 1) / S3_T;
  int since = 0;
  for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const size_t lo = tile * S3_T;
    const int cnt = (int)min((size_t)S3_T, rows - lo);
    for (int i = tid; i < M * DOM; i += S3_THREADS) cur[i] = 0;
    __syncthreads();
    // pass 1: payload rows into shared memory, bucket sizes
#pragma unroll
    for (int u = 0; u < S3_T / S3_THREADS; u++) {
      const int row = tid + u * S3_THREADS;
      if (row < cnt) {
        const size_t r = lo + row;
        pay[row * 3 + 0] = make_float4(1.f, c.num[0][r], c.num[1][r], c.num[2][r]);
        pay[row * 3 + 1] = make_float4(c.num[3][r], c.num[4][r], c.num[5][r], c.num[6][r]);
        pay[row * 3 + 2] = make_float4(c.num[7][r], c.num[8][r], c.num[9][r], 0.f);
#pragma unroll
        for (int k = 0; k < M; k++) atomicAdd(&cur[k * DOM + c.cat[k][r]], 1u);
      }
    }
    __syncthreads();
    // exclusive scan of the bucket sizes, one warp per column
    {
      const int w = tid >> 5, lane = tid & 31;
      if (w < M) {
        unsigned run = 0;
        for (int base = 0; base < DOM; base += 32) {
          const int i = base + lane;
          const unsigned v = i < DOM ? cur[w * DOM + i] : 0;
          unsigned x = v;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
          }
          if (i < DOM) {
            off[w * DOM + i] = run + x - v;
            cur[w * DOM + i] = run + x - v;
          }
          run += __shfl_sync(0xffffffffu, x, 31);
        }
      }
    }
    __syncthreads();
    // pass 2: row ids in bucket order
#pragma unroll
    for (int u = 0; u < S3_T / S3_THREADS; u++) {
      const int row = tid + u * S3_THREADS;
      if (row < cnt) {
        const size_t r = lo + row;
#pragma unroll
        for (int k = 0; k < M; k++) {
          const unsigned pos = atomicAdd(&cur[k * DOM + c.cat[k][r]], 1u);
          ids[k * S3_T + pos] = (unsigned short)row;
        }
      }
    }
    __syncthreads();
    // reduce: thread b = (column, key) adds the payload rows of its bucket
    if (tid < M * DOM) {
      const int k = tid / DOM;
      const unsigned b0 = off[tid], b1 = cur[tid];
      for (unsigned i = b0; i < b1; i++) {
        const int row = ids[k * S3_T + i];
        const float4 p0 = pay[row * 3], p1 = pay[row * 3 + 1], p2 = pay[row * 3 + 2];
        acc[0].x += p0.x; acc[0].y += p0.y; acc[0].z += p0.z; acc[0].w += p0.w;
        acc[1].x += p1.x; acc[1].y += p1.y; acc[1].z += p1.z; acc[1].w += p1.w;
        acc[2].x += p2.x; acc[2].y += p2.y; acc[2].z += p2.z; acc[2].w += p2.w;
      }
    }
    __syncthreads();
    if (++since >= flush_tiles || tile + gridDim.x >= ntiles) {
      since = 0;
      if (tid < M * DOM) {
        double *dst = sums + (size_t)tid * P;
        const float v[12] = {acc[0].x, acc[0].y, acc[0].z, acc[0].w, acc[1].x, acc[1].y, acc[1].z, acc[1].w, acc[2].x, acc[2].y, acc[2].z, acc[2].w};
#pragma unroll
        for (int j = 0; j < 12; j++)
          if (v[j] != 0.f) atomicAdd(dst + j, (double)v[j]);
        acc[0] = acc[1] = acc[2] = make_float4(0, 0, 0, 0);
      }
    }
  }
}

// ------------------------------------------------------------------ S1: warp-private fp32 tables, plain RMW
constexpr int S1_WARPS = 4, S1_ROWS = 10;  // 10 rows x 3 lanes (float4 each) per step
__global__ void __launch_bounds__(S1_WARPS * 32) s1_kernel(Cols c, size_t rows, double *sums, int chunk_rows) {
  extern __shared__ float4 smem_f4[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 *tbl = smem_f4 + (size_t)warp * M * DOM * 3;  // [M][DOM][3 float4]
  const int g = lane / 3, q = lane % 3;                // row of the step, quarter of the payload
  const bool lane_on = lane < 3 * S1_ROWS;
  const size_t gw = (size_t)blockIdx.x * S1_WARPS + warp, nw = (size_t)gridDim.x * S1_WARPS;
  const size_t nchunks = (rows + chunk_rows - 1) / chunk_rows;
  for (int i = lane; i < M * DOM * 3; i += 32) tbl[i] = make_float4(0, 0, 0, 0);
  __syncwarp();
  for (size_t ch = gw; ch < nchunks; ch += nw) {
    const size_t lo = ch * chunk_rows, hi = min(rows, lo + (size_t)chunk_rows);
    for (size_t r0 = lo; r0 < hi; r0 += S1_ROWS) {
      const size_t r = r0 + g;
      const bool live = lane_on && r < hi;
      float4 pay = make_float4(0, 0, 0, 0);
      if (live) {
        if (q == 0) pay = make_float4(1.f, c.num[0][r], c.num[1][r], c.num[2][r]);
        else if (q == 1) pay = make_float4(c.num[3][r], c.num[4][r], c.num[5][r], c.num[6][r]);
        else pay = make_float4(c.num[7][r], c.num[8][r], c.num[9][r], 0.f);
      }
      int key[M];
#pragma unroll
      for (int k = 0; k < M; k++) key[k] = live ? c.cat[k][r] : (0x40000000 | lane);
      unsigned rank[M];
      unsigned multi = 0;
#pragma unroll
      for (int k = 0; k < M; k++) {
        const unsigned mask = __match_any_sync(0xffffffffu, key[k]);
        rank[k] = __popc(mask & ((1u << lane) - 1)) / 3;
        multi |= rank[k];
      }
      // round 0 for all columns: loads first, then adds, then stores (independent tables)
      float4 cur[M];
#pragma unroll
      for (int k = 0; k < M; k++)
        if (live && rank[k] == 0) cur[k] = tbl[(k * DOM + key[k]) * 3 + q];
#pragma unroll
      for (int k = 0; k < M; k++)
        if (live && rank[k] == 0) {
          cur[k].x += pay.x; cur[k].y += pay.y; cur[k].z += pay.z; cur[k].w += pay.w;
          tbl[(k * DOM + key[k]) * 3 + q] = cur[k];
        }
      // later rounds: rows of the step that share a key with an earlier row
      if (__any_sync(0xffffffffu, multi != 0)) {
        for (unsigned round = 1; __any_sync(0xffffffffu, multi >= round); round++) {
          __syncwarp();
#pragma unroll
          for (int k = 0; k < M; k++)
            if (live && rank[k] == round) {
              float4 v = tbl[(k * DOM + key[k]) * 3 + q];
              v.x += pay.x; v.y += pay.y; v.z += pay.z; v.w += pay.w;
              tbl[(k * DOM + key[k]) * 3 + q] = v;
            }
          multi = 0;
#pragma unroll
          for (int k = 0; k < M; k++) multi = max(multi, rank[k]);
        }
      }
      __syncwarp();
    }
    // fold into the fp64 totals
    for (int i = lane; i < M * DOM * 3; i += 32) {
      const float4 v = tbl[i];
      if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) {
        atomicAdd(sums + (size_t)i * 4 + 0, (double)v.x);
        atomicAdd(sums + (size_t)i * 4 + 1, (double)v.y);
        atomicAdd(sums + (size_t)i * 4 + 2, (double)v.z);
        atomicAdd(sums + (size_t)i * 4 + 3, (double)v.w);
        tbl[i] = make_float4(0, 0, 0, 0);
      }
    }
    __syncwarp();
  }
}

int main(int argc, char **argv) {
  const size_t rows = argc > 1 ? strtoull(argv[1], 0, 10) : 64000000ull;
  int sms;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  Cols c;
  for (int k = 0; k < M; k++) {
    int *p; CK(cudaMalloc(&p, rows * 4));
    gen<<<sms * 8, 256>>>(p, nullptr, rows, 1000 + k);
    c.cat[k] = p;
  }
  for (int k = 0; k < N; k++) {
    float *p; CK(cudaMalloc(&p, rows * 4));
    gen<<<sms * 8, 256>>>(nullptr, p, rows, 2000 + k);
    c.num[k] = p;
  }
  CK(cudaDeviceSynchronize());
  const size_t pair_cells = (size_t)NPAIR * DOM * DOM, sum_cells = (size_t)M * DOM * P;
  unsigned long long *pairs0, *pairs1;
  float *sums0; double *sums1;
  CK(cudaMalloc(&pairs0, pair_cells * 8)); CK(cudaMalloc(&pairs1, pair_cells * 8));
  CK(cudaMalloc(&sums0, sum_cells * 4)); CK(cudaMalloc(&sums1, sum_cells * 8));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto report = [&](const char *name, float ms) {
    printf("%-58s %9.3f ms  %8.2f G rows/s\n", name, ms, rows / ms / 1e6);
    fflush(stdout);
  };
  float ms;
  // P0
  for (int rep = 0; rep < 2; rep++) {
    CK(cudaMemset(pairs0, 0, pair_cells * 8));
    CK(cudaEventRecord(e0));
    p0_kernel<<<sms * 8, 256>>>(c, rows, pairs0);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  }
  CK(cudaEventElapsedTime(&ms, e0, e1)); report("P0 pairs: red.global.add.u64 (45 per row)", ms);
  // P1
  CK(cudaFuncSetAttribute(p1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P1_TABLES * DOM * DOM * 4));
  for (int rep = 0; rep < 2; rep++) {
    CK(cudaMemset(pairs1, 0, pair_cells * 8));
    CK(cudaEventRecord(e0));
    p1_kernel<<<(sms / P1_ROLES) * P1_ROLES, 512, P1_TABLES * DOM * DOM * 4>>>(c, rows, pairs1, 1 << 20);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  }
  CK(cudaEventElapsedTime(&ms, e0, e1)); report("P1 pairs: shared u32 atomics (5 tables/CTA, 9 roles)", ms);
  // P2 (with and without match)
  CK(cudaFuncSetAttribute(p2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_TABLES * DOM * DOM * 2));
  CK(cudaFuncSetAttribute(p2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_TABLES * DOM * DOM * 2));
  for (int rep = 0; rep < 2; rep++) {
    CK(cudaMemset(pairs1, 0, pair_cells * 8));
    CK(cudaEventRecord(e0));
    p2_kernel<false><<<(sms / P2_ROLES) * P2_ROLES, P2_TABLES * 32, P2_TABLES * DOM * DOM * 2>>>(c, rows, pairs1, 65024);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  }
  CK(cudaEventElapsedTime(&ms, e0, e1)); report("P2' pairs: plain u16 RMW, NO match (wrong; prices match)", ms);
  for (int rep = 0; rep < 2; rep++) {
    CK(cudaMemset(pairs1, 0, pair_cells * 8));
    CK(cudaEventRecord(e0));
    p2_kernel<true><<<(sms / P2_ROLES) * P2_ROLES, P2_TABLES * 32, P2_TABLES * DOM * DOM * 2>>>(c, rows, pairs1, 65024);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  }
  CK(cudaEventElapsedTime(&ms, e0, e1)); report("P2 pairs: match.any + plain u16 RMW (9 tables/CTA, 5 roles)", ms);
  {
    std::vector<unsigned long long> a(pair_cells), b(pair_cells);
    CK(cudaMemcpy(a.data(), pairs0, pair_cells * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), pairs1, pair_cells * 8, cudaMemcpyDeviceToHost));
    size_t bad = 0; unsigned long long tot = 0;
    for (size_t i = 0; i < pair_cells; i++) { bad += a[i] != b[i]; tot += b[i]; }
    printf("   P2 vs P0: %zu cells differ of %zu, total count %llu (expect %llu)\n", bad, pair_cells, tot, (unsigned long long)rows * NPAIR);
  }
  // P3
  auto run_p3 = [&](int threads, cudaStream_t st) {
    const int smem = P2_TABLES * DOM * DOM * 2, grid = (sms / P2_ROLES) * P2_ROLES;
    if (threads == 512) p3_kernel<512><<<grid, 512, smem, st>>>(c, rows, pairs1, 65024);
    else p3_kernel<1024><<<grid, 1024, smem, st>>>(c, rows, pairs1, 65024);
  };
  CK(cudaFuncSetAttribute(p3_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_TABLES * DOM * DOM * 2));
  CK(cudaFuncSetAttribute(p3_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_TABLES * DOM * DOM * 2));
  for (int threads : {512, 1024}) {
    for (int rep = 0; rep < 2; rep++) {
      CK(cudaMemset(pairs1, 0, pair_cells * 8));
      CK(cudaEventRecord(e0));
      run_p3(threads, 0);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    }
    CK(cudaEventElapsedTime(&ms, e0, e1));
    char nm[128]; snprintf(nm, sizeof nm, "P3 pairs: shared atomics on packed u16 (9 tables/CTA, 5 roles), %d thr", threads);
    report(nm, ms);
    std::vector<unsigned long long> a(pair_cells), b(pair_cells);
    CK(cudaMemcpy(a.data(), pairs0, pair_cells * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), pairs1, pair_cells * 8, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t i = 0; i < pair_cells; i++) bad += a[i] != b[i];
    printf("   P3 vs P0: %zu cells differ\n", bad);
  }
  // S0b: per-CTA slabs
  float *slabs; const int s0b_grid = sms * 4;
  CK(cudaMalloc(&slabs, (size_t)s0b_grid * sum_cells * 4));
  for (int rep = 0; rep < 2; rep++) {
    CK(cudaMemset(slabs, 0, (size_t)s0b_grid * sum_cells * 4));
    CK(cudaEventRecord(e0));
    s0b_kernel<<<s0b_grid, 256>>>(c, rows, slabs);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  }
  CK(cudaEventElapsedTime(&ms, e0, e1)); report("S0b sums: red.global.add.v4.f32 (30 per row), per-CTA slabs", ms);
  // P3 and S0b concurrently on two streams
  {
    cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
    cudaEvent_t j; CK(cudaEventCreate(&j));
    for (int rep = 0; rep < 2; rep++) {
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0, s1));
      CK(cudaStreamWaitEvent(s2, e0));
      run_p3(512, s1);
      s0b_kernel<<<s0b_grid, 256, 0, s2>>>(c, rows, slabs);
      CK(cudaEventRecord(j, s2));
      CK(cudaStreamWaitEvent(s1, j));
      CK(cudaEventRecord(e1, s1)); CK(cudaEventSynchronize(e1));
    }
    CK(cudaEventElapsedTime(&ms, e0, e1)); report("P3(512) || S0b on two streams (pairs + sums together)", ms);
  }
  // S2
  CK(cudaFuncSetAttribute(s2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, M * DOM * 3 * 16));
  for (int per_sm : {1, 2, 4}) {
    for (int rep = 0; rep < 2; rep++) {
      CK(cudaMemset(sums1, 0, sum_cells * 8));
      CK(cudaEventRecord(e0));
      s2_kernel<<<sms * per_sm, M * 32, M * DOM * 3 * 16>>>(c, rows, sums1, 16380);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    }
    CK(cudaEventElapsedTime(&ms, e0, e1));
    char nm[128]; snprintf(nm, sizeof nm, "S2 sums: CTA-shared tables, owner warp per column, %d CTA/SM", per_sm);
    report(nm, ms);
  }
  std::vector<double> s2_host(sum_cells);
  CK(cudaMemcpy(s2_host.data(), sums1, sum_cells * 8, cudaMemcpyDeviceToHost));
  // S3
  const int s3_smem = S3_T * 48 + M * S3_T * 2 + 2 * M * DOM * 4;
  CK(cudaFuncSetAttribute(s3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, s3_smem));
  for (int rep = 0; rep < 2; rep++) {
    CK(cudaMemset(sums1, 0, sum_cells * 8));
    CK(cudaEventRecord(e0));
    s3_kernel<<<sms, S3_THREADS, s3_smem>>>(c, rows, sums1, 10);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  }
  CK(cudaEventElapsedTime(&ms, e0, e1)); report("S3 sums: tile counting sort + bucket-owner threads (no float atomics)", ms);
  std::vector<double> s3_host(sum_cells);
  CK(cudaMemcpy(s3_host.data(), sums1, sum_cells * 8, cudaMemcpyDeviceToHost));
  {
    double w3 = 0;
    for (size_t i = 0; i < sum_cells; i++) if (s2_host[i] != 0) w3 = fmax(w3, fabs(s3_host[i] - s2_host[i]) / fabs(s2_host[i]));
    printf("   S3 vs S2: max rel diff %.3g\n", w3);
  }
  // S0
  for (int rep = 0; rep < 2; rep++) {
    CK(cudaMemset(sums0, 0, sum_cells * 4));
    CK(cudaEventRecord(e0));
    s0_kernel<<<sms * 8, 256>>>(c, rows, sums0);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  }
  CK(cudaEventElapsedTime(&ms, e0, e1)); report("S0 sums: red.global.add.v4.f32 (30 per row), one shared table", ms);
  // S1
  const int s1_smem = S1_WARPS * M * DOM * 3 * 16;
  CK(cudaFuncSetAttribute(s1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, s1_smem));
  for (int chunk : {16380, 65520}) {
    for (int rep = 0; rep < 2; rep++) {
      CK(cudaMemset(sums1, 0, sum_cells * 8));
      CK(cudaEventRecord(e0));
      s1_kernel<<<sms, S1_WARPS * 32, s1_smem>>>(c, rows, sums1, chunk);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    }
    CK(cudaEventElapsedTime(&ms, e0, e1));
    char nm[128]; snprintf(nm, sizeof nm, "S1 sums: warp-private fp32 tables, LDS.128 RMW, chunk %d", chunk);
    report(nm, ms);
  }
  {
    std::vector<float> a(sum_cells); std::vector<double> b(sum_cells);
    CK(cudaMemcpy(a.data(), sums0, sum_cells * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), sums1, sum_cells * 8, cudaMemcpyDeviceToHost));
    double worst = 0; size_t bad_counts = 0;
    for (size_t i = 0; i < sum_cells; i++) {
      if (i % P == 0) { bad_counts += (double)a[i] != b[i] && a[i] < 16777216.f; continue; }
      if (b[i] != 0) worst = fmax(worst, fabs(a[i] - b[i]) / fabs(b[i]));
    }
    double w2 = 0;
    for (size_t i = 0; i < sum_cells; i++) if (b[i] != 0) w2 = fmax(w2, fabs(s2_host[i] - b[i]) / fabs(b[i]));
    printf("   S2 vs S1: max rel diff %.3g\n", w2);
    printf("   S1 vs S0: max rel diff of sums %.3g (S0 is fp32: expect ~1e-4), count cells differing %zu\n", worst, bad_counts);
  }
  return 0;
}
