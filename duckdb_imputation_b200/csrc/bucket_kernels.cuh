// bucket_kernels.cuh -- kernel family K3, the per-key payload sums without float atomics.
//
// Replaces `payload[0] += 1; payload[k+1] += x_k` per (row, categorical column) of
// Triple::SumNoLift (sum_no_lift.cpp:158-189): for every categorical column c and key, the count
// and the sums of the n numeric columns over the rows with c = key (lin_cat / quad_num_cat).
//
// Scatter-adding a 1+n float payload per (row, column) costs 30 L2 vector reductions per row on
// C3 and L2 executes ~200 G of them per second whatever their width (csrc/micro/cat_probe.cu).
// This kernel does no float atomics at all.  Per tile of T rows, in shared memory:
//   1. every thread copies the payload rows [1, x_0..x_{n-1}] of its rows into the tile and counts
//      the bucket sizes (bucket = (column, key); one integer ATOMS per row and column);
//   2. block-wide exclusive scan of the bucket sizes;
//   3. every thread writes its row ids in bucket order (a second ATOMS per row and column);
//   4. the payload rows of a bucket are added up by its owner lanes (one per 16-byte quad of the payload)
//      with plain LDS.128 / FADD; the result is added to this CTA's fp32 slab (global, L2-resident) with
//      a plain load/store -- a bucket has exactly one set of owner lanes in the CTA;
//   5. every `fold_tiles` tiles (<= ~32 K rows: bounds every fp32 run) the slab is folded into the
//      fp64 / u64 state.
// Measured on the B200 (C3 shape, profiles/r01_cat_probe.txt): 14.5 G rows/s against 6.6 G rows/s for
// the L2 reductions.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "slab_kernels.cuh"
#include "state_layout.h"

namespace cfb {

constexpr int kBucketThreads = 1024;
constexpr int kBucketMaxDom = 4096;  // slots x sum of the domains: two u32 tables of this size live in shared memory

struct BucketArgs {
  ScanCols cols;
  unsigned long long n_rows;
  int m, total_dom;
  int n_groups;    // GROUP BY slots: bucket = (slot, column, key); n_groups * total_dom <= kBucketMaxDom
  long long F, U;  // per-slot strides of the f64 / u64 state
  int tile_rows;   // rows per tile (<= 65536: row ids are 16-bit)
  int fold_tiles;  // fold the slab into the state every this many tiles of a CTA
  int lo[kMaxCat], dom[kMaxCat], cat_off[kMaxCat + 1];
  long long numcat_base;
  float *slab;  // [gridDim.x][n_groups * total_dom * P], all zero on entry and on exit
  double *f64;
  unsigned long long *u64;
  int *err;
};

// dynamic shared memory: payload tile, row ids in bucket order, bucket starts and cursors
__host__ __device__ inline size_t bucket_smem_bytes(int n, int m, int buckets, int tile_rows) {
  return (size_t)tile_rows * pad4(1 + n) * 4 + (size_t)m * tile_rows * 2 + (size_t)2 * buckets * 4;
}

template <int N>
__global__ void __launch_bounds__(kBucketThreads, 1) bucket_sum_kernel(const __grid_constant__ BucketArgs a) {
  extern __shared__ float4 bucket_smem[];
  constexpr int P = pad4(1 + N), Q = P / 4;
  const int T = a.tile_rows, m = a.m, D = a.n_groups * a.total_dom, tid = threadIdx.x;  // D buckets
  float4 *pay = bucket_smem;                                             // [T][Q]
  unsigned short *ids = reinterpret_cast<unsigned short *>(pay + (size_t)T * Q);  // [m * T] row ids, bucket order
  unsigned *off = reinterpret_cast<unsigned *>(ids + (size_t)m * T);     // [D] first entry of the bucket
  unsigned *cur = off + D;                                               // [D] sizes, then cursors
  __shared__ unsigned warp_tot[32];
  float *slab = a.slab + (size_t)blockIdx.x * D * P;

  const unsigned long long n_tiles = (a.n_rows + T - 1) / T;
  int since_fold = 0;
  for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const unsigned long long lo = tile * T;
    const int cnt = (int)min((unsigned long long)T, a.n_rows - lo);
    for (int i = tid; i < D; i += kBucketThreads) cur[i] = 0;
    __syncthreads();
    // ---- 1. payload rows + bucket sizes
    bool bad = false;
    for (int row = tid; row < cnt; row += kBucketThreads) {
      const unsigned long long r = lo + row;
      int gbase = 0;
      if (a.cols.group) {
        const int g = a.cols.group[r];
        if (g < 0 || g >= a.n_groups) {  // < 0: filtered row
          if (g > 0) atomicExch(a.err, 2);
          continue;
        }
        gbase = g * a.total_dom;
      }
      float v[P];
      v[0] = 1.f;
#pragma unroll
      for (int k = 0; k < N; k++) v[1 + k] = a.cols.num[k][r];
#pragma unroll
      for (int k = 1 + N; k < P; k++) v[k] = 0.f;
#pragma unroll
      for (int q = 0; q < Q; q++) pay[(size_t)row * Q + q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      for (int c0 = 0; c0 < m; c0 += 4) {  // keys four columns at a time: the loads are independent
        unsigned s[4];
#pragma unroll
        for (int e = 0; e < 4; e++) s[e] = c0 + e < m ? (unsigned)(a.cols.cat[c0 + e][r] - a.lo[c0 + e]) : 0u;
#pragma unroll
        for (int e = 0; e < 4; e++)
          if (c0 + e < m) {
            if (s[e] < (unsigned)a.dom[c0 + e]) atomicAdd(&cur[gbase + a.cat_off[c0 + e] + s[e]], 1u);
            else bad = true;
          }
      }
    }
    if (bad) atomicExch(a.err, 1);  // a key outside the declared domain: the scan reports CFB_ERR_DOMAIN
    __syncthreads();
    // ---- 2. exclusive scan of the bucket sizes (4 consecutive buckets per thread)
    {
      unsigned x[4], sum = 0;
#pragma unroll
      for (int e = 0; e < 4; e++) {
        x[e] = 4 * tid + e < D ? cur[4 * tid + e] : 0u;
        sum += x[e];
      }
      unsigned incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= o) incl += y;
      }
      if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
      __syncthreads();
      if (tid < 32) {
        const unsigned w = warp_tot[tid];
        unsigned wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned y = __shfl_up_sync(0xffffffffu, wi, o);
          if (tid >= o) wi += y;
        }
        warp_tot[tid] = wi - w;  // exclusive prefix of the warp totals
      }
      __syncthreads();
      unsigned run = warp_tot[tid >> 5] + incl - sum;
#pragma unroll
      for (int e = 0; e < 4; e++)
        if (4 * tid + e < D) {
          off[4 * tid + e] = run;
          cur[4 * tid + e] = run;
          run += x[e];
        }
    }
    __syncthreads();
    // ---- 3. row ids in bucket order
    for (int row = tid; row < cnt; row += kBucketThreads) {
      const unsigned long long r = lo + row;
      int gbase = 0;
      if (a.cols.group) {
        const int g = a.cols.group[r];
        if (g < 0 || g >= a.n_groups) continue;
        gbase = g * a.total_dom;
      }
      for (int c0 = 0; c0 < m; c0 += 4) {
        unsigned s[4];
#pragma unroll
        for (int e = 0; e < 4; e++) s[e] = c0 + e < m ? (unsigned)(a.cols.cat[c0 + e][r] - a.lo[c0 + e]) : 0u;
#pragma unroll
        for (int e = 0; e < 4; e++)
          if (c0 + e < m && s[e] < (unsigned)a.dom[c0 + e]) ids[atomicAdd(&cur[gbase + a.cat_off[c0 + e] + s[e]], 1u)] = (unsigned short)row;
      }
    }
    __syncthreads();
    // ---- 4. every bucket is added up by its Q owner lanes (one quad of the payload each: the lanes of a
    //         bucket read one payload row as contiguous 16-byte pieces) and goes to the CTA's slab
    {
      constexpr int BPW = 32 / Q;  // buckets per warp and pass
      const int lane = tid & 31, bw = lane / Q, q = lane - bw * Q;
      for (int base = (tid >> 5) * BPW; base < D; base += (kBucketThreads / 32) * BPW) {
        const int b = base + bw;
        if (bw >= BPW || b >= D) continue;
        const unsigned b0 = off[b], b1 = cur[b];
        if (b0 == b1) continue;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (unsigned i = b0; i < b1; i++) {
          const float4 w = pay[(size_t)ids[i] * Q + q];
          acc.x += w.x, acc.y += w.y, acc.z += w.z, acc.w += w.w;
        }
        float4 *dst = reinterpret_cast<float4 *>(slab + (size_t)b * P) + q;
        float4 w = __ldcg(dst);
        w.x += acc.x, w.y += acc.y, w.z += acc.z, w.w += acc.w;
        __stcg(dst, w);
      }
    }
    __syncthreads();
    // ---- 5. fold the slab into the fp64 / u64 state
    if (++since_fold >= a.fold_tiles || tile + gridDim.x >= n_tiles) {
      since_fold = 0;
      for (int i = tid; i < D * P; i += kBucketThreads) {
        const float v = __ldcg(slab + i);
        if (v == 0.f) continue;
        __stcg(slab + i, 0.f);
        const int b = i / P, j = i % P, g = b / a.total_dom, key = b % a.total_dom;
        if (j == 0)
          red_u64(a.u64 + g * a.U + 1 + key, (unsigned long long)v);
        else if (j <= N)
          atomicAdd(a.f64 + g * a.F + a.numcat_base + (long long)(j - 1) * a.total_dom + key, (double)v);
      }
      __syncthreads();
    }
  }
}

}  // namespace cfb
