// sigma_kernels.cuh -- SURVEY 8 f4: the trainers' moment ("sigma") matrix and their solves, on the device.
//
//   sigma_*_kernel      assemble the p x p fp64 moment matrix of the one-hot expanded design
//                       [1 | numeric columns | one column per (categorical column, key)] from a cofactor: what
//                       build_sigma_matrix does on the host (ML/utils.cpp:176-310), as gathers / scatters over the
//                       cofactor's flat arrays (a finalized result) or over the dense device state of a context.
//   standardize_*       standardize_sigma (ML/utils.cpp:580-599).
//   ridge_bgd_kernel    the reference's batch gradient descent with Barzilai-Borwein steps and backtracking line
//                       search (ML/regression.cpp:30-110, :157-238), ONE persistent cooperative kernel: the matrix
//                       stays in L1 / L2, every CTA owns a band of rows of Sigma * theta (a warp per row), the
//                       O(p) vector algebra and all decisions are replayed identically by every CTA from the same
//                       bits, so one grid barrier per matrix-vector product is the only communication.
//   lda_*, chol_*       within-class covariance, shrinkage, blocked Cholesky solve (lda.cpp:196-316).
//
// Everything here is fp64: p^2 values of a few MB, reused thousands of times -- latency- and L2-bound, not HBM-bound.
#pragma once
#include <cstdint>

namespace cfb {

// ---------------------------------------------------------------------------------------------- assembly
// One-hot layout: column k's keys occupy sigma indices col_base[k] .. (col_base[k] < 0: the column is left out,
// i.e. the categorical label of LDA); a key's index inside its column is its rank among the column's keys
// (`rank_of_entry`, -1 = dropped by drop_first).
struct SigmaFromResult {
  int p, n, m;
  long long N;
  const double *lin, *quad;         // [n], [n(n+1)/2] packed upper triangle
  long long total_keys;
  const long long *cat_offsets;     // [m + 1]
  const int *entry_index;           // [total_keys]: sigma index of that (column, key) or -1
  const long long *cat_counts;      // [total_keys]
  const double *numcat;             // [n * total_keys]
  long long n_pairs;                // all pair entries
  const int *pair_a, *pair_b;       // [n_pairs]: sigma indices of the two keys (or -1)
  const long long *pair_counts;     // [n_pairs]
};

__global__ void sigma_from_result_kernel(SigmaFromResult a, double *sigma) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int p = a.p, n = a.n;
  // numeric block (utils.cpp:180-199)
  for (long long t = tid; t < (long long)(n + 1) * (n + 1); t += stride) {
    const int r = (int)(t / (n + 1)), c = (int)(t % (n + 1));
    double v;
    if (r == 0 && c == 0) v = (double)a.N;
    else if (r == 0) v = a.lin[c - 1];
    else if (c == 0) v = a.lin[r - 1];
    else {
      const int i = min(r, c) - 1, j = max(r, c) - 1;
      v = a.quad[(long long)i * n - (long long)i * (i + 1) / 2 + j];
    }
    sigma[(long long)r * p + c] = v;
  }
  // key counts: first row, first column, diagonal (utils.cpp:206-229)
  for (long long t = tid; t < a.total_keys; t += stride) {
    const int i = a.entry_index[t];
    if (i < 0) continue;
    const double v = (double)a.cat_counts[t];
    sigma[i] = v;
    sigma[(long long)i * p] = v;
    sigma[(long long)i * p + i] = v;
  }
  // per-key numeric sums (utils.cpp:231-255)
  for (long long t = tid; t < (long long)n * a.total_keys; t += stride) {
    const int num = (int)(t / a.total_keys) + 1;
    const int i = a.entry_index[t % a.total_keys];
    if (i < 0) continue;
    const double v = a.numcat[t];
    sigma[(long long)i * p + num] = v;
    sigma[(long long)num * p + i] = v;
  }
  // pair counts (utils.cpp:259-309); the (k, k) lists only hold the diagonal, written above with the same value
  for (long long t = tid; t < a.n_pairs; t += stride) {
    const int x = a.pair_a[t], y = a.pair_b[t];
    if (x < 0 || y < 0) continue;
    const double v = (double)a.pair_counts[t];
    sigma[(long long)x * p + y] = v;
    sigma[(long long)y * p + x] = v;
  }
}

// The same matrix straight from a context's dense device state (state_layout.h): cell_index[cat_off[k] + slot] is
// the sigma index of that (column, key) or -1 (key absent, column left out, dropped first key).
struct SigmaFromState {
  int p, n, m;
  const double *f64;                // this group's [lin | quad | numcat]
  const unsigned long long *u64;    // this group's [N | counts | pairs]
  const int *cell_index;            // [total_dom]
  long long total_dom, numcat_base, pair_base;
};

__global__ void sigma_from_state_kernel(SigmaFromState a, double *sigma) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int p = a.p, n = a.n;
  for (long long t = tid; t < (long long)(n + 1) * (n + 1); t += stride) {
    const int r = (int)(t / (n + 1)), c = (int)(t % (n + 1));
    double v;
    if (r == 0 && c == 0) v = (double)a.u64[0];
    else if (r == 0) v = a.f64[c - 1];
    else if (c == 0) v = a.f64[r - 1];
    else {
      const int i = min(r, c) - 1, j = max(r, c) - 1;
      v = a.f64[n + (long long)i * n - (long long)i * (i + 1) / 2 + j];
    }
    sigma[(long long)r * p + c] = v;
  }
  for (long long t = tid; t < a.total_dom; t += stride) {
    const int i = a.cell_index[t];
    if (i < 0) continue;
    const double v = (double)a.u64[1 + t];
    sigma[i] = v;
    sigma[(long long)i * p] = v;
    sigma[(long long)i * p + i] = v;
  }
  for (long long t = tid; t < (long long)n * a.total_dom; t += stride) {
    const int num = (int)(t / a.total_dom) + 1;
    const int i = a.cell_index[t % a.total_dom];
    if (i < 0) continue;
    const double v = a.f64[a.numcat_base + t];
    sigma[(long long)i * p + num] = v;
    sigma[(long long)num * p + i] = v;
  }
}

// one launch per column pair (k < l): the dom_k x dom_l block of pair counts
__global__ void sigma_pairs_from_state_kernel(const unsigned long long *pairs, const int *cell_k, const int *cell_l, int dom_k,
                                              int dom_l, int p, double *sigma) {
  const long long cells = (long long)dom_k * dom_l;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < cells; t += (long long)gridDim.x * blockDim.x) {
    const int x = cell_k[t / dom_l], y = cell_l[t % dom_l];
    if (x < 0 || y < 0) continue;
    const double v = (double)pairs[t];
    sigma[(long long)x * p + y] = v;
    sigma[(long long)y * p + x] = v;
  }
}

// ------------------------------------------------------------------------------------------ standardize
// standardize_sigma (utils.cpp:580-599) in three steps: moments from row 0 and the diagonal; the (i, j >= 1) block
// (reads row 0 / column 0, which step three clears).
__global__ void standardize_moments_kernel(const double *sigma, int p, double *means, double *stds) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p; i += gridDim.x * blockDim.x) {
    const double mu = sigma[i] / sigma[0];
    means[i] = mu;
    const double q = sigma[i] / sigma[0];
    stds[i] = sqrt(sigma[(long long)i * p + i] / sigma[0] - q * q);
  }
}
__global__ void standardize_block_kernel(double *sigma, int p, const double *means, const double *stds) {
  const long long cells = (long long)p * p;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < cells; t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t / p), j = (int)(t % p);
    if (i == 0 || j == 0) continue;
    sigma[t] = (sigma[t] - means[i] * sigma[j] - means[j] * sigma[i] + sigma[0] * means[j] * means[i]) / (stds[i] * stds[j]);
  }
}
__global__ void standardize_clear_kernel(double *sigma, int p) {
  for (int i = 1 + blockIdx.x * blockDim.x + threadIdx.x; i < p; i += gridDim.x * blockDim.x) {
    sigma[i] = 0.0;
    sigma[(long long)i * p] = 0.0;
  }
}

// ------------------------------------------------------------------------------- ridge regression by BGD
constexpr int kBgdThreads = 1024;

struct BgdArgs {
  const double *sigma;  // p x p, row-major (symmetric)
  int p, label;         // label: index of the label's coefficient (numeric column + 1)
  float step_size, lambda;
  int max_iterations;
  int rows_per_cta, warps_per_row;  // the band of rows a CTA owns; warps (1, 2, .. 32) sharing one row
  int small_groups;                 // > 0: ONE CTA, a thread per row and group of columns (small matrices)
  double *v[2];         // two [p] buffers for Sigma * theta (alternating)
  unsigned *barrier;    // zeroed before the launch
  double *theta_out;    // [p]
  double *scalars_out;  // [0] iterations  [1] last error  [2] theta^T Sigma theta / N of the final theta  [3] matrix-vector products
};

__device__ __forceinline__ double ld_cg_f64(const double *p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

// K sums at once; every thread of every CTA returns the same bits: fixed strides, fixed trees (shuffle tree inside a
// warp, one shared-memory exchange, the same tree over the 32 warp results).  `red` holds K * 32 doubles.
template <int K>
__device__ __forceinline__ void bgd_block_sums(double (&x)[K], double *red) {
  const int lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < K; k++)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x[k] += __shfl_xor_sync(0xffffffffu, x[k], o);
  __syncthreads();  // `red` may still be read from the previous sums
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < K; k++) red[k * 32 + (threadIdx.x >> 5)] = x[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; k++) {
    double v = lane < nwarps ? red[k * 32 + lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    x[k] = v;
  }
}
__device__ __forceinline__ double bgd_block_sum(double x, double *red) {
  double a[1] = {x};
  bgd_block_sums<1>(a, red);
  return a[0];
}

struct BgdGrid {
  unsigned *counter;
  unsigned target;
  __device__ void sync() {
    __syncthreads();
    if (gridDim.x == 1) return;  // one CTA: the block barrier orders its global writes for its own threads
    if (threadIdx.x == 0) {
      target += gridDim.x;
      __threadfence();
      atomicAdd(counter, 1u);
      unsigned seen;
      do {
        asm volatile("ld.global.acquire.gpu.u32 %0, [%1];" : "=r"(seen) : "l"(counter));
      } while ((int)(seen - target) < 0);
      __threadfence();
    }
    __syncthreads();
  }
};

// vs = Sigma * theta (in shared memory, the same bits in every CTA), returns theta^T vs.  A CTA owns rows_per_cta
// consecutive rows -- the same ones every time, so a band that fits L1 is served from there after the first product
// -- and puts warps_per_row warps on each (fixed split, fixed order of the partial sums: the result does not depend
// on timing).  With more than one CTA the bands are exchanged through global memory and one grid barrier.
__device__ double bgd_matvec(const BgdArgs &a, const double *theta, double *vs, int &flip, BgdGrid &grid, double *red) {
  if (a.small_groups > 0) {
    // A matrix of a few hundred KB on one CTA: a warp per row spends most of its instructions on reducing 32 lanes
    // to one number; here thread (g, i) walks row i's entries j = g, g + G, .. -- read as sigma[j][i], the matrix is
    // symmetric, so a warp's loads are one contiguous run -- and the G partial sums meet in shared memory.
    const int p = a.p, G = a.small_groups, width = (p + 31) & ~31;
    double *part = red + 96;  // [G][p]
    const int g = threadIdx.x / width, i = threadIdx.x - g * width;
    if (g < G && i < p) {
      double acc = 0.0, acc2 = 0.0;  // two chains: a DFMA waits ~8 cycles for the one before it
      int j = g;
#pragma unroll 4
      for (; j + G < p; j += 2 * G) {
        acc = fma(a.sigma[(long long)j * p + i], theta[j], acc);
        acc2 = fma(a.sigma[(long long)(j + G) * p + i], theta[j + G], acc2);
      }
      if (j < p) acc = fma(a.sigma[(long long)j * p + i], theta[j], acc);
      part[g * p + i] = acc + acc2;
    }
    __syncthreads();
    double part_sum = 0.0;
    for (int r = threadIdx.x; r < p; r += blockDim.x) {
      double sum = 0.0;
      for (int q = 0; q < G; q++) sum += part[q * p + r];
      vs[r] = sum;
      part_sum = fma(theta[r], sum, part_sum);
    }
    return bgd_block_sum(part_sum, red);  // (its first barrier also publishes vs)
  }
  double *v = a.v[flip];
  flip ^= 1;
  const bool alone = gridDim.x == 1;
  const int p = a.p, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int W = a.warps_per_row, part = warp % W, rows_at_once = (int)(blockDim.x >> 5) / W;
  const int row0 = blockIdx.x * a.rows_per_cta;
  for (int r = warp / W; r - warp / W < a.rows_per_cta; r += rows_at_once) {  // uniform trip count: barriers inside
    const int row = row0 + r;
    const bool live = r < a.rows_per_cta && row < p;
    double acc = 0.0;
    if (live) {
      const double *s = a.sigma + (long long)row * p;
#pragma unroll 8
      for (int j = part * 32 + lane; j < p; j += 32 * W) acc = fma(s[j], theta[j], acc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    }
    if (W > 1) {
      __syncthreads();
      if (lane == 0) red[warp] = acc;
      __syncthreads();
      if (live && part == 0 && lane == 0) {
        acc = 0.0;
        for (int w = 0; w < W; w++) acc += red[warp + w];
      }
    }
    if (live && part == 0 && lane == 0) {
      if (alone) vs[row] = acc;
      else v[row] = acc;
    }
  }
  grid.sync();
  if (!alone) {
    for (int i = threadIdx.x; i < p; i += blockDim.x) vs[i] = ld_cg_f64(v + i);
    __syncthreads();
  }
  double part_sum = 0.0;
  for (int i = threadIdx.x; i < p; i += blockDim.x) part_sum = fma(theta[i], vs[i], part_sum);
  return bgd_block_sum(part_sum, red);
}

// Dynamic shared memory: 6 * p doubles (theta, prev_theta, grad, prev_grad, update, Sigma * theta) + 96 doubles
// (+ small_groups * p doubles for the one-CTA path).  blockDim.x is 1024, or 512 for the one-CTA path.
__global__ void __launch_bounds__(kBgdThreads, 1) ridge_bgd_kernel(BgdArgs a) {
  extern __shared__ double bgd_smem[];
  const int p = a.p, label = a.label, tid = threadIdx.x;
  double *theta = bgd_smem, *prev_theta = theta + p, *grad = prev_theta + p, *prev_grad = grad + p, *update = prev_grad + p;
  double *vs = update + p, *red = vs + p;
  BgdGrid grid{a.barrier, 0u};
  int flip = 0;
  const double count = a.sigma[0];
  float step = a.step_size;  // the reference keeps step_size and lambda as FLOAT (regression.cpp:121-122)
  const float lambda = a.lambda;
  for (int i = tid; i < p; i += blockDim.x) {
    theta[i] = i == label ? -1.0 : 0.0;
    prev_theta[i] = theta[i];
    grad[i] = prev_grad[i] = update[i] = 0.0;
  }
  __syncthreads();
  double quad_form = bgd_matvec(a, theta, vs, flip, grid, red);
  // compute_gradient (regression.cpp:30-46)
  if (count != 0.0)
    for (int i = tid; i < p; i += blockDim.x) grad[i] = i == label ? 0.0 : vs[i] / count;
  __syncthreads();
  // compute_error (:48-77) from theta^T Sigma theta and SUM_{i >= 1} theta_i^2
  auto error_of = [&](double qf, double sq_norm) { return count == 0.0 ? 0.0 : (qf / count + lambda * (sq_norm - 1.0)) / 2; };
  double two[2] = {0.0, 0.0};
  for (int i = tid; i < p; i += blockDim.x) {
    const double upd = i == 0 ? grad[0] : grad[i] + lambda * theta[i];
    two[0] += upd * upd;
    if (i >= 1) two[1] += theta[i] * theta[i];
  }
  bgd_block_sums<2>(two, red);
  double gradient_norm = two[0] - (double)lambda * lambda;  // label correction (:180)
  const double first_gradient_norm = sqrt(gradient_norm);
  double prev_error = error_of(quad_form, two[1]);
  double error = prev_error;
  int iterations = 1, backtracks = 0;
  do {
    two[0] = two[1] = 0.0;
    for (int i = tid; i < p; i += blockDim.x) {
      const double upd = i == 0 ? grad[0] : grad[i] + lambda * theta[i];
      update[i] = upd;
      two[0] += upd * upd;
      prev_theta[i] = theta[i];
      prev_grad[i] = grad[i];
      const double t = i == label ? -1.0 : theta[i] - step * upd;
      theta[i] = t;
      if (i >= 1) two[1] += t * t;
    }
    bgd_block_sums<2>(two, red);
    gradient_norm = two[0] - (double)lambda * lambda;
    double dparam_norm = step * sqrt(two[0]);
    quad_form = bgd_matvec(a, theta, vs, flip, grid, red);
    error = error_of(quad_form, two[1]);
    int bt = 0;
    while (error > prev_error - (step / 2) * gradient_norm && bt < 500) {
      step /= 2;
      two[0] = two[1] = 0.0;
      for (int i = tid; i < p; i += blockDim.x) {
        const double newp = prev_theta[i] - step * update[i];
        const double dp = theta[i] - newp;
        two[0] += dp * dp;
        const double t = i == label ? -1.0 : newp;
        theta[i] = t;
        if (i >= 1) two[1] += t * t;
      }
      bgd_block_sums<2>(two, red);
      dparam_norm = sqrt(two[0]);
      quad_form = bgd_matvec(a, theta, vs, flip, grid, red);
      error = error_of(quad_form, two[1]);
      bt++;
    }
    backtracks += bt;
    gradient_norm = sqrt(gradient_norm);
    if (dparam_norm < 1e-20 || gradient_norm / (first_gradient_norm + 0.001) < 1e-8) break;
    // compute_gradient of the accepted theta (its Sigma * theta is still in vs), compute_step_size (:79-107)
    double three[3] = {0.0, 0.0, 0.0};
    for (int i = tid; i < p; i += blockDim.x) {
      const double g = count != 0.0 ? (i == label ? 0.0 : vs[i] / count) : grad[i];
      grad[i] = g;
      const double pd = theta[i] - prev_theta[i], gd = g - prev_grad[i];
      three[0] += pd * pd;
      three[1] += gd * gd;
      three[2] += pd * gd;
    }
    bgd_block_sums<3>(three, red);
    const double dss = three[0], gss = three[1], dgs = three[2];
    if (dgs != 0.0 && gss != 0.0) {
      const double ts = dss / dgs, tm = dgs / gss;
      if (!(tm < 0.0 || ts < 0.0)) step = (float)((tm / ts > 0.5) ? tm : ts - 0.5 * tm);
    }
    prev_error = error;
    iterations++;
  } while (iterations < a.max_iterations);
  if (blockIdx.x == 0) {
    for (int i = tid; i < p; i += blockDim.x) a.theta_out[i] = theta[i];
    if (tid == 0) {
      a.scalars_out[0] = (double)iterations;
      a.scalars_out[1] = error;
      a.scalars_out[2] = count != 0.0 ? quad_form / count : 0.0;  // the variance of :245-256 (theta[label] = -1)
      a.scalars_out[3] = (double)(1 + iterations + backtracks);
    }
  }
}


// ------------------------------------------------------------------------------------------------- LDA
// Class sums straight from the dense state (build_sum_vector, lda.cpp:58-144): sums[c][0] = rows of class c,
// sums[c][1 + i] = SUM x_i over the class, sums[c][index of (column, key)] = rows of the class with that key.
struct LdaSumsFromState {
  int p, n, n_classes;
  const int *class_cell;            // [n_classes]: slot of the class key in the label column
  const double *f64;
  const unsigned long long *u64;
  long long total_dom, numcat_base, label_off;  // label_off = cat_off[label]
};
__global__ void lda_sums_from_state_kernel(LdaSumsFromState a, double *sums) {
  const long long cells = (long long)a.n_classes * (a.n + 1);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < cells; t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t / (a.n + 1)), j = (int)(t % (a.n + 1));
    const long long cell = a.label_off + a.class_cell[c];
    sums[(long long)c * a.p + j] = j == 0 ? (double)a.u64[1 + cell] : a.f64[a.numcat_base + (long long)(j - 1) * a.total_dom + cell];
  }
}
// pair block of (label column, other column): label_first = the label is the pair's first column
__global__ void lda_pair_sums_from_state_kernel(const unsigned long long *pairs, const int *class_cell, int n_classes,
                                                const int *cell_other, int dom_label, int dom_other, int label_first, int p,
                                                double *sums) {
  const long long cells = (long long)n_classes * dom_other;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < cells; t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t / dom_other), so = (int)(t % dom_other);
    const int x = cell_other[so];
    if (x < 0) continue;
    const long long at = label_first ? (long long)class_cell[c] * dom_other + so : (long long)so * dom_label + class_cell[c];
    sums[(long long)c * p + x] = (double)pairs[at];
  }
}

// sums[c][j] = (sums[c][j] - means[j] * sums[c][0]) / stds[j], j >= 1   (lda.cpp:205-212)
__global__ void lda_standardize_sums_kernel(double *sums, int n_classes, int p, const double *means, const double *stds) {
  const long long cells = (long long)n_classes * p;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < cells; t += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(t % p);
    if (j == 0) continue;
    sums[t] = (sums[t] - means[j] * sums[t - j]) / stds[j];
  }
}

// S[j][k] = sigma[j+1][k+1] - SUM_c sums[c][j+1] * sums[c][k+1] / sums[c][0]    (lda.cpp:217-249), q = p - 1;
// rhs[c][j] = sums[c][j+1] / sums[c][0] (the class means).  The subtraction runs class by class, as the reference's.
__global__ void lda_within_kernel(const double *sigma, const double *sums, int n_classes, int p, double *S, double *rhs) {
  const int q = p - 1;
  const long long cells = (long long)q * q;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < cells; t += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(t / q), k = (int)(t % q);
    double v = sigma[(long long)(j + 1) * p + (k + 1)];
    for (int c = 0; c < n_classes; c++) {
      const double *s = sums + (long long)c * p;
      v -= (s[j + 1] * s[k + 1]) / s[0];
    }
    S[t] = v;
    if (k == 0)
      for (int c = 0; c < n_classes; c++) rhs[(long long)c * q + j] = sums[(long long)c * p + j + 1] / sums[(long long)c * p];
  }
}

// mu = trace(S) / q ; S = (S * (1 - shrinkage) + [diag] shrinkage * mu) / N     (lda.cpp:255-275); one CTA computes mu
__global__ void lda_trace_kernel(const double *S, int q, double *mu_out) {
  __shared__ double red[32];
  double part = 0.0;
  for (int j = threadIdx.x; j < q; j += blockDim.x) part += S[(long long)j * q + j];
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += red[w];
    *mu_out = s / (float)q;
  }
}
__global__ void lda_shrink_kernel(double *S, int q, float shrinkage, const double *mu, double count) {
  const long long cells = (long long)q * q;
  const float keep = 1 - shrinkage;  // FLOAT arithmetic, as the reference's `(1-shrinkage)`
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < cells; t += (long long)gridDim.x * blockDim.x) {
    double v = S[t] * keep;
    if (t / q == t % q) v += shrinkage * *mu;
    S[t] = v / count;
  }
}

// Blocked right-looking Cholesky S = L L^T (lower triangle, in place, row-major), 64-column panels.
constexpr int kCholNb = 64;
// factor the diagonal block [k, k + nb): one CTA, the block in shared memory
__global__ void __launch_bounds__(256) chol_diag_kernel(double *S, int q, int k, int nb, int *not_spd) {
  __shared__ double a[kCholNb][kCholNb + 1];
  for (int t = threadIdx.x; t < nb * nb; t += blockDim.x) a[t / nb][t % nb] = S[(long long)(k + t / nb) * q + k + t % nb];
  __syncthreads();
  for (int j = 0; j < nb; j++) {
    const double d = a[j][j];
    if (!(d > 0.0)) {
      if (threadIdx.x == 0) *not_spd = 1;
      return;
    }
    const double r = sqrt(d);
    __syncthreads();
    for (int i = j + threadIdx.x; i < nb; i += blockDim.x) a[i][j] = i == j ? r : a[i][j] / r;
    __syncthreads();
    const int rem = nb - j - 1;
    for (int t = threadIdx.x; t < rem * rem; t += blockDim.x) {
      const int i = j + 1 + t / rem, c = j + 1 + t % rem;
      if (c <= i) a[i][c] -= a[i][j] * a[c][j];
    }
    __syncthreads();
  }
  for (int t = threadIdx.x; t < nb * nb; t += blockDim.x) {
    const int i = t / nb, c = t % nb;
    S[(long long)(k + i) * q + k + c] = c <= i ? a[i][c] : 0.0;
  }
}
// panel: rows i >= k + nb, L[i][k..k+nb) = S[i][k..k+nb) * L_kk^{-T}; one thread per row, L_kk in shared memory
__global__ void __launch_bounds__(128) chol_panel_kernel(double *S, int q, int k, int nb) {
  __shared__ double l[kCholNb][kCholNb + 1];
  for (int t = threadIdx.x; t < nb * nb; t += blockDim.x) l[t / nb][t % nb] = S[(long long)(k + t / nb) * q + k + t % nb];
  __syncthreads();
  const int i = k + nb + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= q) return;
  double *row = S + (long long)i * q + k;
  double x[kCholNb];
#pragma unroll 1
  for (int c = 0; c < nb; c++) {
    double v = row[c];
    for (int t = 0; t < c; t++) v -= x[t] * l[c][t];
    x[c] = v / l[c][c];
  }
  for (int c = 0; c < nb; c++) row[c] = x[c];
}
// trailing update: S[i][j] -= SUM_t L[i][k+t] * L[j][k+t] for i >= j >= k + nb; 32 x 32 tiles
__global__ void __launch_bounds__(256) chol_update_kernel(double *S, int q, int k, int nb) {
  __shared__ double li[32][kCholNb + 1], lj[32][kCholNb + 1];
  const int base = k + nb;
  const int ti = blockIdx.y, tj = blockIdx.x;
  if (tj > ti) return;
  const int i0 = base + ti * 32, j0 = base + tj * 32;
  for (int t = threadIdx.x; t < 32 * nb; t += blockDim.x) {
    const int r = t / nb, c = t % nb;
    li[r][c] = i0 + r < q ? S[(long long)(i0 + r) * q + k + c] : 0.0;
    lj[r][c] = j0 + r < q ? S[(long long)(j0 + r) * q + k + c] : 0.0;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 32 * 32; t += blockDim.x) {
    const int r = t / 32, c = t % 32;
    const int i = i0 + r, j = j0 + c;
    if (i >= q || j >= q || j > i) continue;
    double acc = 0.0;
    for (int u = 0; u < nb; u++) acc = fma(li[r][u], lj[c][u], acc);
    S[(long long)i * q + j] -= acc;
  }
}
// solve L L^T x = b for every right-hand side: one CTA per class, b (then x) in shared memory
__global__ void __launch_bounds__(256) chol_solve_kernel(const double *L, int q, double *rhs) {
  extern __shared__ double x[];
  __shared__ double pivot;
  double *b = rhs + (long long)blockIdx.x * q;
  for (int i = threadIdx.x; i < q; i += blockDim.x) x[i] = b[i];
  __syncthreads();
  for (int j = 0; j < q; j++) {  // forward: L y = b, column-oriented
    if (threadIdx.x == 0) pivot = x[j] = x[j] / L[(long long)j * q + j];
    __syncthreads();
    const double xj = pivot;
    for (int i = j + 1 + threadIdx.x; i < q; i += blockDim.x) x[i] -= L[(long long)i * q + j] * xj;
    __syncthreads();
  }
  for (int j = q - 1; j >= 0; j--) {  // backward: L^T x = y, row j of L is column j of L^T
    if (threadIdx.x == 0) pivot = x[j] = x[j] / L[(long long)j * q + j];
    __syncthreads();
    const double xj = pivot;
    for (int i = threadIdx.x; i < j; i += blockDim.x) x[i] -= L[(long long)j * q + i] * xj;
    __syncthreads();
  }
  for (int i = threadIdx.x; i < q; i += blockDim.x) b[i] = x[i];
}
// intercept[c] = -0.5 * mean_c . x_c + log(count_c / N)   (lda.cpp:316-320); coef[c][j] /= stds[j+1] when normalized
__global__ void lda_intercept_kernel(const double *sums, const double *x, int n_classes, int p, double count, const double *stds,
                                     double *coef, double *intercept) {
  __shared__ double red[32];
  const int c = blockIdx.x, q = p - 1;
  const double *s = sums + (long long)c * p;
  double part = 0.0;
  for (int j = threadIdx.x; j < q; j += blockDim.x) part += (s[j + 1] / s[0]) * x[(long long)c * q + j];
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double dot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) dot += red[w];
    intercept[c] = dot * (-0.5) + log(s[0] / count);
  }
  for (int j = threadIdx.x; j < q; j += blockDim.x) coef[(long long)c * q + j] = stds ? x[(long long)c * q + j] / stds[j + 1] : x[(long long)c * q + j];
}

}  // namespace cfb
