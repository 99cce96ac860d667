// sigma_train.cuh -- host half of SURVEY 8 f4 (included at the end of cofactor_b200.cu): the cfb_sigma_* entry points.
// The one-hot layout is the reference's (n_cols_1hot_expansion, ML/utils.cpp:522-576): per categorical column the keys
// that occur, ordered as the reference orders them (ascending as uint64 of the sign-extended key, utils.cpp:541-556:
// negative keys sort after the positive ones), `cat_vars_idxs` = running offsets.
#pragma once
#include "sigma_kernels.cuh"

struct cfb_sigma {
  int device = 0, p = 0, n = 0, m = 0, label_cat = -1, drop_first = 0;
  long long N = 0;
  double *d_sigma = nullptr;           // p x p
  double *d_sums = nullptr;            // LDA: [n_classes][p] class sums
  int n_classes = 0;
  std::vector<int64_t> cat_array;      // every column's keys (the label column's too), reference order
  std::vector<int32_t> cat_idxs;       // [m + 1]
  cudaStream_t stream = nullptr;
  double *d_bgd = nullptr;             // scratch of cfb_sigma_linreg_train: [2p v | p theta | 4 scalars | barrier], kept
};

namespace {

struct SigmaScratch {  // device buffers freed on scope exit
  std::vector<void *> ptrs;
  ~SigmaScratch() {
    for (void *p : ptrs) cudaFree(p);
  }
  template <class T>
  cudaError_t upload(const std::vector<T> &v, const T **out, cudaStream_t s) {
    void *d = nullptr;
    cudaError_t e = cudaMalloc(&d, std::max<size_t>(1, v.size()) * sizeof(T));
    if (e != cudaSuccess) return e;
    ptrs.push_back(d);
    *out = (const T *)d;
    return v.empty() ? cudaSuccess : cudaMemcpyAsync(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
  }
  template <class T>
  cudaError_t alloc(size_t count, T **out) {
    void *d = nullptr;
    cudaError_t e = cudaMalloc(&d, std::max<size_t>(1, count) * sizeof(T));
    if (e != cudaSuccess) return e;
    ptrs.push_back(d);
    *out = (T *)d;
    return cudaSuccess;
  }
};

inline bool ref_key_less(int32_t a, int32_t b) { return (uint64_t)(int64_t)a < (uint64_t)(int64_t)b; }

// per column: the keys in reference order -> cat_array / cat_idxs; rank[k][j] = position of the column's j-th
// ascending key in that order minus the dropped first key (-1 = dropped)
void one_hot_layout(int m, const std::vector<std::vector<int32_t>> &keys_asc, int drop_first, cfb_sigma *s,
                    std::vector<std::vector<int>> &rank) {
  s->cat_array.clear();
  s->cat_idxs.assign(m + 1, 0);
  rank.assign(m, {});
  for (int k = 0; k < m; k++) {
    std::vector<int> order(keys_asc[k].size());
    for (size_t j = 0; j < order.size(); j++) order[j] = (int)j;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return ref_key_less(keys_asc[k][a], keys_asc[k][b]); });
    rank[k].assign(order.size(), -1);
    int pos = 0;
    for (size_t r = 0; r < order.size(); r++) {
      if (drop_first && r == 0) continue;
      rank[k][order[r]] = pos++;
      s->cat_array.push_back(keys_asc[k][order[r]]);
    }
    s->cat_idxs[k + 1] = s->cat_idxs[k] + pos;
  }
}

// sigma index of (column k, rank r): after the numeric block, columns in order, the label column left out
inline int sigma_index(const cfb_sigma *s, int k, int r) {
  if (r < 0 || k == s->label_cat) return -1;
  const int skipped = (s->label_cat >= 0 && k > s->label_cat) ? s->cat_idxs[s->label_cat + 1] - s->cat_idxs[s->label_cat] : 0;
  return 1 + s->n + s->cat_idxs[k] + r - skipped;
}

int sigma_alloc(cfb_sigma *s) {
  const int label_keys = s->label_cat >= 0 ? s->cat_idxs[s->label_cat + 1] - s->cat_idxs[s->label_cat] : 0;
  s->p = 1 + s->n + s->cat_idxs[s->m] - label_keys;
  s->n_classes = label_keys;
  if (s->p > 4600) return fail(CFB_ERR_DOMAIN, "one-hot expansion of %d columns: more than 4600", s->p);
  CU(cudaMalloc(&s->d_sigma, (size_t)s->p * s->p * sizeof(double)));
  CU(cudaMemsetAsync(s->d_sigma, 0, (size_t)s->p * s->p * sizeof(double), s->stream));
  if (s->n_classes) {
    CU(cudaMalloc(&s->d_sums, (size_t)s->n_classes * s->p * sizeof(double)));
    CU(cudaMemsetAsync(s->d_sums, 0, (size_t)s->n_classes * s->p * sizeof(double), s->stream));
  }
  return CFB_OK;
}

inline int sigma_grid(long long work) { return (int)std::max<long long>(1, std::min<long long>(1184, (work + 255) / 256)); }

}  // namespace

extern "C" void cfb_sigma_destroy(cfb_sigma *s) {
  if (!s) return;
  cudaSetDevice(s->device);
  if (s->d_sigma) cudaFree(s->d_sigma);
  if (s->d_sums) cudaFree(s->d_sums);
  if (s->d_bgd) cudaFree(s->d_bgd);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
}

extern "C" int cfb_sigma_from_result(int device, const cfb_result *res, int label_cat, int drop_first, cfb_sigma **out) {
  NvtxRange nvtx_range("cfb_sigma_from_result");
  if (!res || !out) return fail(CFB_ERR_INVALID, "NULL argument");
  *out = nullptr;
  if (res->kind != CFB_TRIPLE) return fail(CFB_ERR_INVALID, "the sigma matrix needs the full ring (CFB_TRIPLE)");
  if (label_cat >= res->n_cat) return fail(CFB_ERR_INVALID, "categorical label %d out of range", label_cat);
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) {
    cudaGetLastError();
    return fail(CFB_ERR_NO_DEVICE, "CUDA device %d not available (no CPU fallback)", device);
  }
  CU(cudaSetDevice(device));
  std::unique_ptr<cfb_sigma, void (*)(cfb_sigma *)> s(new cfb_sigma, cfb_sigma_destroy);
  s->device = device;
  s->n = res->n_num;
  s->m = res->n_cat;
  s->label_cat = label_cat < 0 ? -1 : label_cat;
  s->drop_first = drop_first ? 1 : 0;
  s->N = res->N;
  CU(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  const int n = s->n, m = s->m;
  std::vector<std::vector<int32_t>> keys(m);
  for (int k = 0; k < m; k++) keys[k].assign(res->cat_keys + res->cat_offsets[k], res->cat_keys + res->cat_offsets[k + 1]);
  std::vector<std::vector<int>> rank;
  one_hot_layout(m, keys, s->drop_first, s.get(), rank);
  int rc = sigma_alloc(s.get());
  if (rc) return rc;
  const long long tk = res->total_keys;
  std::vector<int> entry_index((size_t)tk, -1);
  for (int k = 0; k < m; k++)
    for (size_t j = 0; j < keys[k].size(); j++) entry_index[(size_t)res->cat_offsets[k] + j] = sigma_index(s.get(), k, rank[k][j]);
  auto find = [&](int k, int32_t key) -> long long {  // entry of (column, key) or -1
    auto it = std::lower_bound(keys[k].begin(), keys[k].end(), key);
    return it != keys[k].end() && *it == key ? res->cat_offsets[k] + (it - keys[k].begin()) : -1;
  };
  const long long n_pairs = res->n_pair_lists ? res->pair_offsets[res->n_pair_lists] : 0;
  std::vector<int> pa((size_t)n_pairs, -1), pb((size_t)n_pairs, -1);
  std::vector<long long> pcounts(res->pair_counts, res->pair_counts + n_pairs);
  std::vector<double> sums((size_t)s->n_classes * s->p, 0.0);
  const int lab = s->label_cat;
  auto class_of = [&](int32_t key) -> int {  // class index = the label key's rank in reference order
    const long long e = find(lab, key);
    return e < 0 ? -1 : rank[lab][(size_t)(e - res->cat_offsets[lab])];
  };
  long long list = 0;
  for (int k = 0; k < m; k++)
    for (int l = k; l < m; l++, list++)
      for (long long t = res->pair_offsets[list]; t < res->pair_offsets[list + 1]; t++) {
        const long long ea = find(k, res->pair_key1[t]), eb = find(l, res->pair_key2[t]);
        if (ea < 0 || eb < 0) continue;
        pa[(size_t)t] = entry_index[(size_t)ea];
        pb[(size_t)t] = entry_index[(size_t)eb];
        if (lab >= 0 && k != l && (k == lab || l == lab)) {  // class sums of the other column's keys (lda.cpp:96-143)
          const int c = class_of(k == lab ? res->pair_key1[t] : res->pair_key2[t]);
          const int other = k == lab ? entry_index[(size_t)eb] : entry_index[(size_t)ea];
          if (c >= 0 && other >= 0) sums[(size_t)c * s->p + other] = (double)res->pair_counts[t];
        }
      }
  if (lab >= 0)
    for (size_t j = 0; j < keys[lab].size(); j++) {
      const int c = rank[lab][j];
      if (c < 0) continue;
      const size_t e = (size_t)res->cat_offsets[lab] + j;
      sums[(size_t)c * s->p] = (double)res->cat_counts[e];
      for (int i = 0; i < n; i++) sums[(size_t)c * s->p + 1 + i] = res->numcat_sums[(size_t)i * tk + e];
    }
  SigmaScratch tmp;
  cfb::SigmaFromResult a{};
  a.p = s->p;
  a.n = n;
  a.m = m;
  a.N = res->N;
  a.total_keys = tk;
  a.n_pairs = n_pairs;
  std::vector<double> lin(res->lin, res->lin + n), quad(res->quad, res->quad + res->n_quad);
  std::vector<double> numcat(res->numcat_sums, res->numcat_sums + (size_t)n * tk);
  std::vector<long long> offs(res->cat_offsets, res->cat_offsets + m + 1), counts(res->cat_counts, res->cat_counts + tk);
  CU(tmp.upload(lin, &a.lin, s->stream));
  CU(tmp.upload(quad, &a.quad, s->stream));
  CU(tmp.upload(offs, &a.cat_offsets, s->stream));
  CU(tmp.upload(entry_index, &a.entry_index, s->stream));
  CU(tmp.upload(counts, &a.cat_counts, s->stream));
  CU(tmp.upload(numcat, &a.numcat, s->stream));
  CU(tmp.upload(pa, &a.pair_a, s->stream));
  CU(tmp.upload(pb, &a.pair_b, s->stream));
  CU(tmp.upload(pcounts, &a.pair_counts, s->stream));
  const long long work = std::max<long long>({(long long)(n + 1) * (n + 1), (long long)n * tk, n_pairs, tk});
  cfb::sigma_from_result_kernel<<<sigma_grid(work), 256, 0, s->stream>>>(a, s->d_sigma);
  g_launches++;
  CU(cudaGetLastError());
  if (s->n_classes)
    CU(cudaMemcpyAsync(s->d_sums, sums.data(), sums.size() * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  *out = s.release();
  return CFB_OK;
}

extern "C" int cfb_sigma_from_ctx(cfb_ctx *c, int group, int label_cat, int drop_first, cfb_sigma **out) {
  NvtxRange nvtx_range("cfb_sigma_from_ctx");
  if (!c || !out) return fail(CFB_ERR_INVALID, "NULL argument");
  *out = nullptr;
  if (c->kind != CFB_TRIPLE) return fail(CFB_ERR_INVALID, "the sigma matrix needs the full ring (CFB_TRIPLE)");
  if (group < 0 || group >= c->G) return fail(CFB_ERR_INVALID, "group %d out of range", group);
  if (label_cat >= c->m) return fail(CFB_ERR_INVALID, "categorical label %d out of range", label_cat);
  if (c->lay.pairs_hashed || any_dict(c)) {  // sparse state: canonical result first, then the same scatter
    cfb_result res;
    int rc = cfb_ctx_finalize(c, group, &res);
    if (rc) return rc;
    rc = cfb_sigma_from_result(c->device, &res, label_cat, drop_first, out);
    cfb_result_free(&res);
    return rc;
  }
  int rc = cfb_ctx_sync(c);
  if (rc) return rc;
  const Layout &L = c->lay;
  const int n = c->n, m = c->m;
  std::unique_ptr<cfb_sigma, void (*)(cfb_sigma *)> s(new cfb_sigma, cfb_sigma_destroy);
  s->device = c->device;
  s->n = n;
  s->m = m;
  s->label_cat = label_cat < 0 ? -1 : label_cat;
  s->drop_first = drop_first ? 1 : 0;
  CU(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  // the only read-back: N and the key counts (which keys occur)
  const long long td = m ? L.total_dom : 0;
  std::vector<unsigned long long> head((size_t)(1 + td));
  const unsigned long long *u64 = c->d_u64 + (long long)group * L.U;
  const double *f64 = c->d_f64 + (long long)group * L.F;
  CU(cudaMemcpyAsync(head.data(), u64, head.size() * 8, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  s->N = (long long)head[0];
  std::vector<std::vector<int32_t>> keys(m);
  std::vector<std::vector<int>> slot_of(m);
  for (int k = 0; k < m; k++)
    for (int sl = 0; sl < L.dom[k]; sl++)
      if (head[(size_t)(1 + L.cat_off[k] + sl)]) {
        keys[k].push_back((int32_t)((long long)L.lo[k] + sl));
        slot_of[k].push_back(sl);
      }
  std::vector<std::vector<int>> rank;
  one_hot_layout(m, keys, s->drop_first, s.get(), rank);
  rc = sigma_alloc(s.get());
  if (rc) return rc;
  std::vector<int> cell_index((size_t)td, -1), class_cell((size_t)s->n_classes, 0);
  for (int k = 0; k < m; k++)
    for (size_t j = 0; j < keys[k].size(); j++) {
      cell_index[(size_t)(L.cat_off[k] + slot_of[k][j])] = sigma_index(s.get(), k, rank[k][j]);
      if (k == s->label_cat && rank[k][j] >= 0) class_cell[(size_t)rank[k][j]] = slot_of[k][j];
    }
  SigmaScratch tmp;
  const int *d_cell = nullptr, *d_class_cell = nullptr;
  CU(tmp.upload(cell_index, &d_cell, s->stream));
  CU(tmp.upload(class_cell, &d_class_cell, s->stream));
  cfb::SigmaFromState a{};
  a.p = s->p;
  a.n = n;
  a.m = m;
  a.f64 = f64;
  a.u64 = u64;
  a.cell_index = d_cell;
  a.total_dom = td;
  a.numcat_base = L.numcat_base;
  a.pair_base = L.pair_base;
  const long long work = std::max<long long>({(long long)(n + 1) * (n + 1), (long long)n * td, td});
  cfb::sigma_from_state_kernel<<<sigma_grid(work), 256, 0, s->stream>>>(a, s->d_sigma);
  g_launches++;
  for (int k = 0; k < m; k++)
    for (int l = k + 1; l < m; l++) {
      const unsigned long long *pairs = u64 + L.pair_base + L.pair_off[k * m + l];
      if (k != s->label_cat && l != s->label_cat) {
        cfb::sigma_pairs_from_state_kernel<<<sigma_grid((long long)L.dom[k] * L.dom[l]), 256, 0, s->stream>>>(
            pairs, d_cell + L.cat_off[k], d_cell + L.cat_off[l], L.dom[k], L.dom[l], s->p, s->d_sigma);
        g_launches++;
      } else if (s->n_classes) {
        const int other = k == s->label_cat ? l : k;
        cfb::lda_pair_sums_from_state_kernel<<<sigma_grid((long long)s->n_classes * L.dom[other]), 256, 0, s->stream>>>(
            pairs, d_class_cell, s->n_classes, d_cell + L.cat_off[other], L.dom[s->label_cat], L.dom[other], k == s->label_cat, s->p,
            s->d_sums);
        g_launches++;
      }
    }
  if (s->n_classes) {
    cfb::LdaSumsFromState b{};
    b.p = s->p;
    b.n = n;
    b.n_classes = s->n_classes;
    b.class_cell = d_class_cell;
    b.f64 = f64;
    b.u64 = u64;
    b.total_dom = td;
    b.numcat_base = L.numcat_base;
    b.label_off = L.cat_off[s->label_cat];
    cfb::lda_sums_from_state_kernel<<<sigma_grid((long long)s->n_classes * (n + 1)), 256, 0, s->stream>>>(b, s->d_sums);
    g_launches++;
  }
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s->stream));
  *out = s.release();
  return CFB_OK;
}

extern "C" int cfb_sigma_shape(const cfb_sigma *s, int32_t *p, int32_t *n_classes, int64_t *n_cat_values) {
  if (!s) return fail(CFB_ERR_INVALID, "sigma is NULL");
  if (p) *p = s->p;
  if (n_classes) *n_classes = s->n_classes;
  if (n_cat_values) *n_cat_values = (int64_t)s->cat_array.size();
  return CFB_OK;
}

extern "C" int cfb_sigma_layout(const cfb_sigma *s, int64_t *cat_array, int32_t *cat_vars_idxs) {
  if (!s) return fail(CFB_ERR_INVALID, "sigma is NULL");
  if (cat_array) std::copy(s->cat_array.begin(), s->cat_array.end(), cat_array);
  if (cat_vars_idxs) std::copy(s->cat_idxs.begin(), s->cat_idxs.end(), cat_vars_idxs);
  return CFB_OK;
}

extern "C" int cfb_sigma_download(const cfb_sigma *s, double *sigma, double *class_sums) {
  if (!s) return fail(CFB_ERR_INVALID, "sigma is NULL");
  CU(cudaSetDevice(s->device));
  if (sigma) CU(cudaMemcpyAsync(sigma, s->d_sigma, (size_t)s->p * s->p * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  if (class_sums && s->n_classes)
    CU(cudaMemcpyAsync(class_sums, s->d_sums, (size_t)s->n_classes * s->p * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return CFB_OK;
}

namespace {
// standardize_sigma into `work` (a copy: the handle keeps the raw moments), means / stds on the device
int sigma_standardized_copy(const cfb_sigma *s, SigmaScratch &tmp, double **work, double **d_means, double **d_stds) {
  const size_t cells = (size_t)s->p * s->p;
  CU(tmp.alloc(cells, work));
  CU(tmp.alloc((size_t)s->p, d_means));
  CU(tmp.alloc((size_t)s->p, d_stds));
  CU(cudaMemcpyAsync(*work, s->d_sigma, cells * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
  cfb::standardize_moments_kernel<<<sigma_grid(s->p), 256, 0, s->stream>>>(*work, s->p, *d_means, *d_stds);
  cfb::standardize_block_kernel<<<sigma_grid((long long)cells), 256, 0, s->stream>>>(*work, s->p, *d_means, *d_stds);
  cfb::standardize_clear_kernel<<<sigma_grid(s->p), 256, 0, s->stream>>>(*work, s->p);
  g_launches += 3;
  CU(cudaGetLastError());
  return CFB_OK;
}
}  // namespace

extern "C" int cfb_sigma_linreg_train(cfb_sigma *s, int label, float step_size, float lambda, int max_iterations, int normalize,
                                      double *coeff, double *means, double *variance, int32_t *iterations, int32_t *products) {
  NvtxRange nvtx_range("cfb_sigma_linreg_train");
  if (!s || !coeff) return fail(CFB_ERR_INVALID, "NULL argument");
  if (s->label_cat >= 0) return fail(CFB_ERR_STATE, "this sigma matrix leaves a categorical label out (LDA): build one with label_cat = -1");
  if (label < 0 || label >= s->n) return fail(CFB_ERR_INVALID, "label %d is not a numeric column (0..%d)", label, s->n - 1);
  CU(cudaSetDevice(s->device));
  const int p = s->p;
  SigmaScratch tmp;
  double *work = s->d_sigma, *d_means = nullptr, *d_stds = nullptr;
  if (normalize) {
    int rc = sigma_standardized_copy(s, tmp, &work, &d_means, &d_stds);
    if (rc) return rc;
  }
  cfb::BgdArgs a{};
  a.sigma = work;
  a.p = p;
  a.label = label + 1;  // index 0 is the intercept (regression.cpp:170)
  a.step_size = step_size;
  a.lambda = lambda;
  a.max_iterations = max_iterations;
  // scratch stays with the handle: a MICE step trains on the same shape again and again
  if (!s->d_bgd) CU(cudaMalloc(&s->d_bgd, ((size_t)3 * p + 4 + 2) * sizeof(double)));
  double *d_v = s->d_bgd, *d_theta = d_v + 2 * (size_t)p, *d_scalars = d_theta + p;
  unsigned *d_barrier = reinterpret_cast<unsigned *>(d_scalars + 4);
  CU(cudaMemsetAsync(d_barrier, 0, sizeof(unsigned), s->stream));
  a.v[0] = d_v;
  a.v[1] = d_v + p;
  a.barrier = d_barrier;
  a.theta_out = d_theta;
  a.scalars_out = d_scalars;
  int sms = 1;
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device));
  // bands of >= 8 rows (a band that fits L1 stays there between products), several warps on a row when the band is
  // shorter than the CTA's 32 warps; a matrix of a few hundred KB is cheaper on ONE CTA than a grid barrier per product
  static const int single_cta_max_p = getenv("CFB_BGD_SINGLE_CTA_MAX_P") ? atoi(getenv("CFB_BGD_SINGLE_CTA_MAX_P")) : 160;
  const int grid = p <= single_cta_max_p ? 1 : std::max(1, std::min(sms, (p + 7) / 8));
  a.rows_per_cta = (p + grid - 1) / grid;
  a.warps_per_row = 1;
  while (a.warps_per_row < 32 && a.rows_per_cta * a.warps_per_row * 2 <= 32) a.warps_per_row *= 2;
  int threads = cfb::kBgdThreads;
  size_t smem = ((size_t)6 * p + 96) * sizeof(double);
  if (grid == 1 && ((p + 31) & ~31) <= 512 && !getenv("CFB_BGD_NO_SMALL")) {  // one CTA of 512 threads, a thread per (row, column group)
    threads = 512;
    a.small_groups = std::max(1, threads / ((p + 31) & ~31));
    smem += (size_t)a.small_groups * p * sizeof(double);
  }
  CU(cudaFuncSetAttribute(cfb::ridge_bgd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 << 10)));
  void *params[] = {&a};
  CU(cudaLaunchCooperativeKernel((const void *)cfb::ridge_bgd_kernel, dim3(grid), dim3(threads), params, smem, s->stream));
  g_launches++;
  std::vector<double> theta((size_t)p), mu, sd;
  double scalars[4];
  CU(cudaMemcpyAsync(theta.data(), d_theta, (size_t)p * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaMemcpyAsync(scalars, d_scalars, sizeof scalars, cudaMemcpyDeviceToHost, s->stream));
  if (normalize) {
    mu.resize((size_t)p);
    sd.resize((size_t)p);
    CU(cudaMemcpyAsync(mu.data(), d_means, (size_t)p * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaMemcpyAsync(sd.data(), d_stds, (size_t)p * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  }
  CU(cudaStreamSynchronize(s->stream));
  if (normalize) {  // back to the raw scale (regression.cpp:262-267)
    const int l = label + 1;
    for (int i = 1; i < p; i++) theta[(size_t)i] = (theta[(size_t)i] / sd[(size_t)i]) * sd[(size_t)l];
    theta[0] = theta[0] * sd[(size_t)l] + mu[(size_t)l];
    if (means) std::copy(mu.begin(), mu.end(), means);
  }
  std::copy(theta.begin(), theta.end(), coeff);
  if (variance) *variance = scalars[2];
  if (iterations) *iterations = (int32_t)scalars[0];
  if (products) *products = (int32_t)scalars[3];
  return CFB_OK;
}

extern "C" int cfb_sigma_lda_train(cfb_sigma *s, float shrinkage, int normalize, double *coef, double *intercept, double *means) {
  NvtxRange nvtx_range("cfb_sigma_lda_train");
  if (!s || !coef || !intercept) return fail(CFB_ERR_INVALID, "NULL argument");
  if (s->label_cat < 0 || s->n_classes == 0) return fail(CFB_ERR_STATE, "this sigma matrix has no categorical label: build one with label_cat >= 0");
  CU(cudaSetDevice(s->device));
  const int p = s->p, q = p - 1, C = s->n_classes;
  if (q < 1) return fail(CFB_ERR_INVALID, "no feature columns");
  SigmaScratch tmp;
  double *work = s->d_sigma, *d_means = nullptr, *d_stds = nullptr, *sums = s->d_sums;
  if (normalize) {
    int rc = sigma_standardized_copy(s, tmp, &work, &d_means, &d_stds);
    if (rc) return rc;
    CU(tmp.alloc((size_t)C * p, &sums));
    CU(cudaMemcpyAsync(sums, s->d_sums, (size_t)C * p * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
    cfb::lda_standardize_sums_kernel<<<sigma_grid((long long)C * p), 256, 0, s->stream>>>(sums, C, p, d_means, d_stds);
    g_launches++;
  }
  double *S = nullptr, *rhs = nullptr, *d_mu = nullptr, *d_coef = nullptr, *d_icpt = nullptr;
  int *d_flag = nullptr;
  CU(tmp.alloc((size_t)q * q, &S));
  CU(tmp.alloc((size_t)C * q, &rhs));
  CU(tmp.alloc((size_t)C * q, &d_coef));
  CU(tmp.alloc((size_t)C, &d_icpt));
  CU(tmp.alloc((size_t)1, &d_mu));
  CU(tmp.alloc((size_t)1, &d_flag));
  CU(cudaMemsetAsync(d_flag, 0, sizeof(int), s->stream));
  cfb::lda_within_kernel<<<sigma_grid((long long)q * q), 256, 0, s->stream>>>(work, sums, C, p, S, rhs);
  cfb::lda_trace_kernel<<<1, 1024, 0, s->stream>>>(S, q, d_mu);
  cfb::lda_shrink_kernel<<<sigma_grid((long long)q * q), 256, 0, s->stream>>>(S, q, shrinkage, d_mu, (double)s->N);
  g_launches += 3;
  for (int k = 0; k < q; k += cfb::kCholNb) {
    const int nb = std::min(cfb::kCholNb, q - k), below = q - k - nb;
    cfb::chol_diag_kernel<<<1, 256, 0, s->stream>>>(S, q, k, nb, d_flag);
    g_launches++;
    if (below > 0) {
      cfb::chol_panel_kernel<<<(below + 127) / 128, 128, 0, s->stream>>>(S, q, k, nb);
      const int tiles = (below + 31) / 32;
      cfb::chol_update_kernel<<<dim3(tiles, tiles), 256, 0, s->stream>>>(S, q, k, nb);
      g_launches += 2;
    }
  }
  const size_t smem = (size_t)q * sizeof(double);
  CU(cudaFuncSetAttribute(cfb::chol_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cfb::chol_solve_kernel<<<C, 256, smem, s->stream>>>(S, q, rhs);
  cfb::lda_intercept_kernel<<<C, 256, 0, s->stream>>>(sums, rhs, C, p, (double)s->N, normalize ? d_stds : nullptr, d_coef, d_icpt);
  g_launches += 2;
  CU(cudaGetLastError());
  int flag = 0;
  CU(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaMemcpyAsync(coef, d_coef, (size_t)C * q * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaMemcpyAsync(intercept, d_icpt, (size_t)C * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  if (normalize && means) CU(cudaMemcpyAsync(means, d_means, (size_t)p * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  if (flag)
    return fail(CFB_ERR_STATE, "the within-class covariance is not positive definite (collinear one-hot columns): use shrinkage > 0");
  return CFB_OK;
}
