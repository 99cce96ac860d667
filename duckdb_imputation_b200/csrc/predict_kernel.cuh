// predict_kernel.cuh -- model scores of rows: the write-back step of a MICE iteration.
//
// Replaces the per-row loops of ML::linreg_impute (ML/regression.cpp:436-506) and LDA_impute
// (ML/lda.cpp:506-577): score_o = bias_o + SUM_i w_num[o][i] x_i + SUM_c w_cat[o][pos_c(key_c)],
// written as the score itself (regression) or the index of the largest score (LDA).
//
// HBM-bound by construction: every input value is read once (4(n+m) bytes per row, coalesced:
// consecutive threads take consecutive rows of the SoA columns) and 4 bytes are written.  The
// model lives in shared memory: bias and weights in fp64 (the reference accumulates in double),
// and per categorical column a dense key -> position map (int32, -1 = unknown key), so a key
// lookup is one LDS instead of the reference's linear search (regression.cpp:471-476).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "state_layout.h"

namespace cfb {

constexpr int kPredictThreads = 256;
constexpr int kPredictMaxOut = 32;

struct PredictArgs {
  ScanCols cols;            // group = row mask (nullable)
  unsigned long long n_rows;
  int n, m, n_out, mode;
  int kb;                   // outputs padded to this many (class-minor weight rows of kb doubles); 1 for n_out = 1
  int vec4;                 // score kernel: all pointers 16-byte aligned -> 4 rows per thread with 128-bit loads
  int total;                // model keys over all columns
  int map_lo[kMaxCat], map_off[kMaxCat], map_len[kMaxCat];  // dense map of column c: [map_off, map_off + map_len) covers keys [map_lo, ..)
  int map_total;
  int col_off[kMaxCat];     // first position of column c in w_cat
  const double *d_model;    // [kb] bias | [n][kb] w_num | [total][kb] w_cat  (padding: weight 0, bias -inf)
  const int *d_map;         // [map_total] position within w_cat rows, or -1
  void *out;
  // stochastic regression (linreg_predict with noise = true, regression.cpp:495-505): score += sigma * N(0, 1), the
  // normal drawn by a counter-based generator from (seed, first_row + row) -- reproducible whatever the chunking
  double noise_sigma;
  unsigned long long noise_seed, noise_first;
};

// Philox-4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3"): a pure function of (key, counter).
__device__ __forceinline__ uint4 philox4x32(unsigned long long key, unsigned long long counter) {
  unsigned k0 = (unsigned)key, k1 = (unsigned)(key >> 32);
  uint4 c = make_uint4((unsigned)counter, (unsigned)(counter >> 32), 0x9E3779B9u, 0xBB67AE85u);
#pragma unroll
  for (int round = 0; round < 10; round++) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}
// One standard normal per (seed, row): Box-Muller on two 53-bit / 32-bit uniforms, the same transform the reference
// applies to libc random() (regression.cpp:496-503).
__device__ __forceinline__ double predict_normal(unsigned long long seed, unsigned long long row) {
  const uint4 r = philox4x32(seed, row);
  const double u1 = ((double)(((unsigned long long)r.x << 21) ^ (r.y >> 11)) + 1.0) * (1.0 / 9007199254740992.0);  // (0, 1]
  const double u2 = (double)r.z * (1.0 / 4294967296.0);                                                          // [0, 1)
  return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

__host__ __device__ inline size_t predict_smem_bytes(int n, int kb, int total, int map_total) {
  return (size_t)kb * (1 + n + total) * 8 + (size_t)map_total * 4;
}

// model -> shared memory (every CTA; persistent grid, so once per CTA)
__device__ __forceinline__ void predict_load_model(const PredictArgs &a, double *smem) {
  const int words = a.kb * (1 + a.n + a.total);
  for (int i = threadIdx.x; i < words; i += blockDim.x) smem[i] = a.d_model[i];
  int *map = reinterpret_cast<int *>(smem + words);
  for (int i = threadIdx.x; i < a.map_total; i += blockDim.x) map[i] = a.d_map[i];
  __syncthreads();
}

// One output (regression): 4 consecutive rows per thread, every column one 128-bit load (columns are
// 16-byte aligned), four independent fp64 accumulators -- the loads of a row step are independent of
// the accumulation, which is what keeps enough bytes in flight to stream from HBM.
__global__ void __launch_bounds__(kPredictThreads, 4) predict_score_kernel(const __grid_constant__ PredictArgs a) {
  extern __shared__ double predict_smem[];
  predict_load_model(a, predict_smem);
  const int n = a.n, m = a.m;
  const double *w_num = predict_smem + 1, *w_cat = w_num + n;
  const int *map = reinterpret_cast<const int *>(w_cat + a.total);
  const double bias = predict_smem[0];
  const unsigned long long n4 = a.vec4 ? a.n_rows / 4 : 0;  // unaligned input: every row takes the scalar path below
  float *out = static_cast<float *>(a.out);
  for (unsigned long long q = (unsigned long long)blockIdx.x * kPredictThreads + threadIdx.x; q < n4;
       q += (unsigned long long)gridDim.x * kPredictThreads) {
    int4 mk = make_int4(1, 1, 1, 1);
    if (a.cols.group) {
      mk = reinterpret_cast<const int4 *>(a.cols.group)[q];
      if (!(mk.x | mk.y | mk.z | mk.w)) continue;  // no cell to fill among these rows
    }
    double acc0 = bias, acc1 = bias, acc2 = bias, acc3 = bias;
#pragma unroll 4
    for (int i = 0; i < n; i++) {
      const float4 x = reinterpret_cast<const float4 *>(a.cols.num[i])[q];
      const double w = w_num[i];
      acc0 += w * (double)x.x, acc1 += w * (double)x.y, acc2 += w * (double)x.z, acc3 += w * (double)x.w;
    }
#pragma unroll 2
    for (int c = 0; c < m; c++) {
      const int4 k = reinterpret_cast<const int4 *>(a.cols.cat[c])[q];
      const int lo = a.map_lo[c], off = a.map_off[c];
      const unsigned len = (unsigned)a.map_len[c];
      const unsigned d0 = (unsigned)(k.x - lo), d1 = (unsigned)(k.y - lo), d2 = (unsigned)(k.z - lo), d3 = (unsigned)(k.w - lo);
      const int p0 = d0 < len ? map[off + d0] : -1, p1 = d1 < len ? map[off + d1] : -1;
      const int p2 = d2 < len ? map[off + d2] : -1, p3 = d3 < len ? map[off + d3] : -1;
      if (p0 >= 0) acc0 += w_cat[p0];
      if (p1 >= 0) acc1 += w_cat[p1];
      if (p2 >= 0) acc2 += w_cat[p2];
      if (p3 >= 0) acc3 += w_cat[p3];
    }
    if (a.noise_sigma != 0.0) {
      const unsigned long long r0 = a.noise_first + 4 * q;
      acc0 += a.noise_sigma * predict_normal(a.noise_seed, r0);
      acc1 += a.noise_sigma * predict_normal(a.noise_seed, r0 + 1);
      acc2 += a.noise_sigma * predict_normal(a.noise_seed, r0 + 2);
      acc3 += a.noise_sigma * predict_normal(a.noise_seed, r0 + 3);
    }
    if (!a.cols.group) {
      reinterpret_cast<float4 *>(out)[q] = make_float4((float)acc0, (float)acc1, (float)acc2, (float)acc3);
    } else {
      if (mk.x) out[4 * q] = (float)acc0;
      if (mk.y) out[4 * q + 1] = (float)acc1;
      if (mk.z) out[4 * q + 2] = (float)acc2;
      if (mk.w) out[4 * q + 3] = (float)acc3;
    }
  }
  // the last n_rows % 4 rows (all rows when the input is not 16-byte aligned)
  for (unsigned long long r = 4 * n4 + (unsigned long long)blockIdx.x * kPredictThreads + threadIdx.x; r < a.n_rows;
       r += (unsigned long long)gridDim.x * kPredictThreads) {
    if (a.cols.group && a.cols.group[r] == 0) continue;
    double acc = bias;
    for (int i = 0; i < n; i++) acc += w_num[i] * (double)a.cols.num[i][r];
    for (int c = 0; c < m; c++) {
      const unsigned d = (unsigned)(a.cols.cat[c][r] - a.map_lo[c]);
      const int pos = d < (unsigned)a.map_len[c] ? map[a.map_off[c] + d] : -1;
      if (pos >= 0) acc += w_cat[pos];
    }
    if (a.noise_sigma != 0.0) acc += a.noise_sigma * predict_normal(a.noise_seed, a.noise_first + r);
    out[r] = (float)acc;
  }
}

// Several outputs (LDA): one row per thread, KB >= n_out accumulators in registers.  The weights are stored
// class-minor ([feature][KB], padded with zeros), so the KB weights of a feature are consecutive: a uniform
// LDS.128 per two classes for a numeric feature, and one contiguous run per thread for a (column, key).
// Result = first index of the largest score (lda.cpp:566-573), or score_0 when mode = SCORE.
// A wide model leaves room for one or two CTAs per SM only; up to 12 outputs (<= 64 registers per thread) the CTA may
// then have 1024 threads, so that one copy of the model serves 32 warps (the kernel waits on its column loads).
constexpr int predict_multi_max_threads(int kb) { return kb <= 12 ? 1024 : kPredictThreads; }
template <int KB>
__global__ void __launch_bounds__(predict_multi_max_threads(KB)) predict_multi_kernel(const __grid_constant__ PredictArgs a) {
  extern __shared__ double predict_smem[];
  predict_load_model(a, predict_smem);
  const int n = a.n, m = a.m;
  const double *bias = predict_smem, *w_num = bias + KB, *w_cat = w_num + (size_t)n * KB;
  const int *map = reinterpret_cast<const int *>(w_cat + (size_t)a.total * KB);
  for (unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; r < a.n_rows;
       r += (unsigned long long)gridDim.x * blockDim.x) {
    if (a.cols.group && a.cols.group[r] == 0) continue;  // not a cell to fill
    double acc[KB];
#pragma unroll
    for (int k = 0; k < KB; k++) acc[k] = bias[k];
#pragma unroll 4
    for (int i = 0; i < n; i++) {
      const double x = (double)a.cols.num[i][r];
      const double2 *w = reinterpret_cast<const double2 *>(w_num + (size_t)i * KB);
#pragma unroll
      for (int k = 0; k < KB / 2; k++) {
        const double2 v = w[k];
        acc[2 * k] += v.x * x;
        acc[2 * k + 1] += v.y * x;
      }
    }
#pragma unroll 2
    for (int c = 0; c < m; c++) {
      const unsigned d = (unsigned)(a.cols.cat[c][r] - a.map_lo[c]);
      const int pos = d < (unsigned)a.map_len[c] ? map[a.map_off[c] + d] : -1;
      if (pos < 0) continue;
      const double2 *w = reinterpret_cast<const double2 *>(w_cat + (size_t)pos * KB);
#pragma unroll
      for (int k = 0; k < KB / 2; k++) {
        const double2 v = w[k];
        acc[2 * k] += v.x;
        acc[2 * k + 1] += v.y;
      }
    }
    double best = acc[0];
    int best_k = 0;
#pragma unroll
    for (int k = 1; k < KB; k++)
      if (acc[k] > best) {  // padded classes carry bias = -inf and never win
        best = acc[k];
        best_k = k;
      }
    if (a.mode == 0) static_cast<float *>(a.out)[r] = (float)acc[0];
    else static_cast<int *>(a.out)[r] = best_k;
  }
}

// ---------------------------------------------------------------------------------------------------------
// nb_predict (ML::nb_impute, ML/naive_bayes.cpp:153-263): per row and class the product of the prior, one Gaussian
// density per numeric column and one probability per categorical column; the LABEL of the first class whose product
// is the largest (class 0 when every product is 0).  fp64 and the reference's own expression, so that the products
// agree with it to rounding.  Model in global memory (read through L1): [K] labels live in `labels`.
struct NbArgs {
  ScanCols cols;  // group = row mask (nullable)
  unsigned long long n_rows;
  int n, m, n_classes, total;
  int map_lo[kMaxCat], map_off[kMaxCat], map_len[kMaxCat];
  const int *d_map;        // dense key -> position in [0, total), or -1
  const int *labels;       // [K]
  const double *prior;     // [K]
  const double *norm;      // [K][n]  1 / sqrt(2 pi (var + 1e-9))
  const double *mean;      // [K][n]
  const double *two_var;   // [K][n]  2 (var + 1e-9)
  const double *cat_prob;  // [K][total]
  int *out;
};

__global__ void __launch_bounds__(kPredictThreads) predict_nb_kernel(const __grid_constant__ NbArgs a) {
  for (unsigned long long r = (unsigned long long)blockIdx.x * kPredictThreads + threadIdx.x; r < a.n_rows;
       r += (unsigned long long)gridDim.x * kPredictThreads) {
    if (a.cols.group && a.cols.group[r] == 0) continue;
    int pos[kMaxCat];
    for (int c = 0; c < a.m; c++) {
      const unsigned d = (unsigned)(a.cols.cat[c][r] - a.map_lo[c]);
      pos[c] = d < (unsigned)a.map_len[c] ? a.d_map[a.map_off[c] + d] : -1;
    }
    int best = 0;
    double max_prob = 0.0;
    for (int k = 0; k < a.n_classes; k++) {
      double prob = a.prior[k];
      for (int j = 0; j < a.n; j++) {
        const double d = (double)a.cols.num[j][r] - a.mean[k * a.n + j];
        prob *= a.norm[k * a.n + j] * exp(-(d * d) / a.two_var[k * a.n + j]);
      }
      for (int c = 0; c < a.m; c++) prob *= pos[c] >= 0 ? a.cat_prob[(size_t)k * a.total + pos[c]] : 0.0;
      if (prob > max_prob) {
        max_prob = prob;
        best = k;
      }
    }
    a.out[r] = a.labels[best];
  }
}

// ---------------------------------------------------------------------------------------------------------
// qda_predict (ML::qda_impute, ML/qda.cpp:338-498): score_k = intercept_k + f^T Q_k f + l_k . f over
// f = [numeric | one-hot] - centre.  The reference multiplies the dense p x p matrix per row and class
// (p = n + number of keys).  Here f = z - c with z sparse (n numeric values and m ones), and the host folds the
// centre into the model: score_k = z^T Q_k z + g_k . z + b_k with g_k = l_k - (Q_k + Q_k^T) c, b_k = intercept_k +
// c^T Q_k c - l_k . c -- (n + m)^2 terms per row and class whatever the number of keys.
struct QdaArgs {
  ScanCols cols;  // group = row mask (nullable)
  unsigned long long n_rows;
  int n, m, n_classes, total, p;  // p = n + total
  int map_lo[kMaxCat], map_off[kMaxCat], map_len[kMaxCat];
  const int *d_map;
  const int *labels;  // [K]
  const double *Q;    // [K][p][p], Q[k][i + j * p] (column-major like the reference's dgemv operand)
  const double *g;    // [K][p]
  const double *b;    // [K]
  int *out;
};

__global__ void __launch_bounds__(kPredictThreads) predict_qda_kernel(const __grid_constant__ QdaArgs a) {
  const int n = a.n, m = a.m, p = a.p;
  for (unsigned long long r = (unsigned long long)blockIdx.x * kPredictThreads + threadIdx.x; r < a.n_rows;
       r += (unsigned long long)gridDim.x * kPredictThreads) {
    if (a.cols.group && a.cols.group[r] == 0) continue;
    double x[32];
    int at[kMaxCat];  // index of the row's one-hot entry of column c in f, or -1
    for (int i = 0; i < n; i++) x[i] = (double)a.cols.num[i][r];
    for (int c = 0; c < m; c++) {
      const unsigned d = (unsigned)(a.cols.cat[c][r] - a.map_lo[c]);
      const int pos = d < (unsigned)a.map_len[c] ? a.d_map[a.map_off[c] + d] : -1;
      at[c] = pos >= 0 ? n + pos : -1;
    }
    int best = 0;
    double max_score = -1.7976931348623157e308;
    for (int k = 0; k < a.n_classes; k++) {
      const double *Q = a.Q + (size_t)k * p * p, *g = a.g + (size_t)k * p;
      double s = a.b[k];
      for (int j = 0; j < n; j++) {  // numeric x numeric, numeric x one-hot, linear
        double col = g[j];
        for (int i = 0; i < n; i++) col += x[i] * Q[i + (size_t)j * p];
        for (int c = 0; c < m; c++)
          if (at[c] >= 0) col += Q[at[c] + (size_t)j * p] + Q[j + (size_t)at[c] * p];
        s += col * x[j];
      }
      for (int c = 0; c < m; c++) {  // one-hot x one-hot, linear
        if (at[c] < 0) continue;
        s += g[at[c]];
        for (int e = 0; e < m; e++)
          if (at[e] >= 0) s += Q[at[c] + (size_t)at[e] * p];
      }
      if (s > max_score) {
        max_score = s;
        best = k;
      }
    }
    a.out[r] = a.labels[best];
  }
}

}  // namespace cfb
