// predict_glue.cpp -- linreg_predict / lda_predict, B200 build: the write-back step of a MICE iteration.
//
//   linreg_predict(params FLOAT[], noise BOOL, normalize BOOL, cols...)  ML::linreg_impute  reference: ML/regression.cpp:397-509
//   lda_predict(params FLOAT[], normalize BOOL, cols...)                 LDA_impute         reference: ML/lda.cpp:421-590
// The parameter lists keep the reference's layout (they are what linreg_train / lda_train emit); the glue
// turns them into a cfb_linear_model -- bias, numeric weights, per-(column,key) weights, with the
// `normalize` centering folded into the bias -- and scores the chunk on the GPU (cfb_predict_host).
//
//   nb_predict(params FLOAT[], normalize BOOL, cols...)                  ML::nb_impute      reference: ML/naive_bayes.cpp:153-263
//   qda_predict(params FLOAT[], normalize BOOL, cols...)                 ML::qda_impute     reference: ML/qda.cpp:338-498
//
// Deliberate differences: a key the model does not know contributes 0 (the reference reads
// past the column's weights, regression.cpp:471-498, or asserts, lda.cpp:528).  `noise = true` adds
// sigma * N(0, 1) like the reference (regression.cpp:495-505), but the normals come from a counter-based
// generator on the device (Philox + Box-Muller keyed by a per-executor seed and the row's position in the
// stream) instead of libc random() seeded from /dev/urandom: same distribution, reproducible with CFB_NOISE_SEED.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#include "../../../include/cofactor_b200.h"
#include "triple_glue.h"

namespace ML {

namespace {

struct Columns {
  std::vector<duckdb::UnifiedVectorFormat> fmt;
  std::vector<const float *> num;
  std::vector<const int32_t *> cat;
  std::vector<const uint32_t *> num_sel, cat_sel;
};

// feature columns start at `first`: FLOAT / DOUBLE -> numeric, INTEGER -> categorical (regression.cpp:410-416)
void Classify(duckdb::DataChunk &args, idx_t first, Columns &c) {
  using namespace duckdb;
  const idx_t rows = args.size();
  c.fmt.resize(args.ColumnCount());
  for (idx_t j = first; j < args.ColumnCount(); j++) {
    args.data[j].ToUnifiedFormat(rows, c.fmt[j]);
    const auto &t = args.data[j].GetType();
    if (t == LogicalType::FLOAT || t == LogicalType::DOUBLE) {
      if (!c.cat.empty()) throw InvalidInputException("numeric columns must precede categorical ones");
      c.num.push_back(UnifiedVectorFormat::GetData<float>(c.fmt[j]));
      c.num_sel.push_back(c.fmt[j].sel->data());
    } else if (t == LogicalType::INTEGER) {
      c.cat.push_back(UnifiedVectorFormat::GetData<int32_t>(c.fmt[j]));
      c.cat_sel.push_back(c.fmt[j].sel->data());
    } else {
      throw InvalidInputException("feature columns must be FLOAT or INTEGER");
    }
  }
}

std::vector<float> Params(duckdb::Vector &v, idx_t rows) {
  using namespace duckdb;
  if (v.GetType().id() != LogicalTypeId::LIST) throw InvalidInputException("the first argument is the parameter list (FLOAT[])");
  UnifiedVectorFormat f;
  v.ToUnifiedFormat(rows, f);
  const list_entry_t e = UnifiedVectorFormat::GetData<list_entry_t>(f)[f.sel->get_index(0)];
  const float *d = FlatVector::GetData<float>(ListVector::GetEntry(v));
  return std::vector<float>(d + e.offset, d + e.offset + e.length);
}

bool Flag(duckdb::Vector &v, idx_t rows) {
  duckdb::UnifiedVectorFormat f;
  v.ToUnifiedFormat(rows, f);
  return duckdb::UnifiedVectorFormat::GetData<uint8_t>(f)[f.sel->get_index(0)] != 0;
}

void Check(int rc) {
  if (rc == CFB_OK) return;
  const std::string msg = cfb_last_error();
  if (rc == CFB_ERR_INVALID || rc == CFB_ERR_DOMAIN) throw duckdb::InvalidInputException(msg);
  throw duckdb::InternalException(msg);
}

int Device() {
  const char *one = getenv("CFB_DEVICE");
  return one ? atoi(one) : 0;
}

struct Model {
  cfb_linear_model m{};
  std::vector<double> bias, w_num, w_cat;
  std::vector<int64_t> offs;
  std::vector<int32_t> keys;
  void Bind() {
    m.bias = bias.data();
    m.w_num = w_num.data();
    m.cat_offsets = offs.data();
    m.cat_keys = keys.data();
    m.w_cat = w_cat.data();
  }
};

// The uploaded model of the calling thread's previous chunk: a query scores thousands of chunks with one
// parameter list, so the device copy is rebuilt only when the list, the flags or the column split change.
struct ModelCache {
  std::vector<float> params;
  int tag = -1;  // function | flags | n | m
  cfb_model *model = nullptr;
  uint64_t noise_seed = 0, rows_done = 0;  // this executor's noise stream and its position
  ~ModelCache() { cfb_model_destroy(model); }
};
thread_local ModelCache t_cache;

cfb_model *Cached(const std::vector<float> &p, int tag) {
  return t_cache.model && t_cache.tag == tag && t_cache.params == p ? t_cache.model : nullptr;
}
cfb_model *Upload(Model &M, const std::vector<float> &p, int tag) {
  cfb_model_destroy(t_cache.model);
  t_cache.model = nullptr;
  M.Bind();
  Check(cfb_model_create(Device(), &M.m, &t_cache.model));
  t_cache.params = p;
  t_cache.tag = tag;
  // a fresh noise stream per uploaded model and executor thread: CFB_NOISE_SEED pins it (tests), else the system's
  // entropy source, as the reference seeds from /dev/urandom (regression.cpp:380-391)
  static std::atomic<uint64_t> executor{0};
  const uint64_t lane = executor.fetch_add(1);
  if (const char *e = getenv("CFB_NOISE_SEED")) t_cache.noise_seed = strtoull(e, nullptr, 10) + 0x9E3779B97F4A7C15ull * lane;
  else t_cache.noise_seed = ((uint64_t)std::random_device{}() << 32) ^ std::random_device{}() ^ lane;
  t_cache.rows_done = 0;
  return t_cache.model;
}

// noise = true: position the executor's stream for this chunk (sigma is the list's last entry, regression.cpp:503)
void ArmNoise(cfb_model *model, bool noise, const std::vector<float> &p, idx_t rows) {
  if (!noise) return;
  Check(cfb_model_set_noise(model, std::fabs((double)p.back()), t_cache.noise_seed, t_cache.rows_done));
  t_cache.rows_done += rows;
}

// The model sorts each column's keys ascending (the trainers emit them in lin_cat order, which is ascending
// already; sorting makes the glue independent of that).
void SortColumns(Model &M) {
  const int K = M.m.n_out;
  const size_t total = M.keys.size();
  std::vector<size_t> perm(total);
  for (size_t i = 0; i < total; i++) perm[i] = i;
  for (int c = 0; c < M.m.n_cat; c++)
    std::sort(perm.begin() + M.offs[c], perm.begin() + M.offs[c + 1], [&](size_t x, size_t y) { return M.keys[x] < M.keys[y]; });
  std::vector<int32_t> k2(total);
  std::vector<double> w2(M.w_cat.size());
  for (size_t i = 0; i < total; i++) {
    k2[i] = M.keys[perm[i]];
    for (int k = 0; k < K; k++) w2[(size_t)k * total + i] = M.w_cat[(size_t)k * total + perm[i]];
  }
  M.keys.swap(k2);
  M.w_cat.swap(w2);
}

}  // namespace

// regression.cpp:397-509.  params = [n_cat | idx_0..idx_{n_cat} | unique keys | intercept | w_num | w_cat |
//                                    (means_num | means_cat when trained with normalize) | sigma]
void linreg_impute(duckdb::DataChunk &args, duckdb::ExpressionState &, duckdb::Vector &result) {
  using namespace duckdb;
  const idx_t rows = args.size();
  if (rows == 0) return;
  if (args.ColumnCount() < 3) throw InvalidInputException("linreg_predict(params, noise, normalize, columns...)");
  const std::vector<float> p = Params(args.data[0], rows);
  const bool noise = Flag(args.data[1], rows);
  const bool normalize = Flag(args.data[2], rows);
  Columns cols;
  Classify(args, 3, cols);
  const size_t n = cols.num.size(), m = cols.cat.size();
  const int tag = (int)(0 | (normalize ? 2 : 0) | (noise ? 4 : 0) | (n << 8) | (m << 16));
  if (cfb_model *hit = Cached(p, tag)) {
    ArmNoise(hit, noise, p, rows);
    Check(cfb_predict_host(hit, cols.num.data(), cols.num_sel.data(), cols.cat.data(), cols.cat_sel.data(), rows,
                           CFB_PREDICT_SCORE, FlatVector::GetData<float>(result)));
    return;
  }
  if (p.empty()) throw InvalidInputException("empty parameter list");
  const size_t n_cat = (size_t)p[0];
  if (n_cat != m) throw InvalidInputException("the parameter list was trained on a different number of categorical columns");
  size_t start = 1 + n_cat, total = 0;  // regression.cpp:428-435
  if (n_cat > 0) {
    if (p.size() <= start) throw InvalidInputException("parameter list too short");
    total = (size_t)p[start];
    start += total + 1;
  }
  const size_t need = start + 1 + n + total + (normalize ? n + total : 0);
  if (p.size() < need) throw InvalidInputException("parameter list too short for these columns");
  Model M;
  M.m.n_num = (int)n;
  M.m.n_cat = (int)m;
  M.m.n_out = 1;
  double bias = p[start];  // intercept
  for (size_t i = 0; i < n; i++) {
    M.w_num.push_back(p[start + 1 + i]);
    if (normalize) bias -= (double)p[start + 1 + i] * (double)p[1 + n + total + start + i];  // (x - mean) * w  (:441-447)
  }
  M.offs.push_back(0);
  for (size_t c = 0; c < m; c++) {
    const size_t b = (size_t)p[1 + c], e = (size_t)p[2 + c];
    if (e < b || e > total) throw InvalidInputException("malformed categorical index in the parameter list");
    for (size_t j = b; j < e; j++) {
      M.keys.push_back((int32_t)p[j + 2 + n_cat]);
      const double w = p[j + start + n + 1];
      M.w_cat.push_back(w);  // the matching key's (1 - mean) * w = w - mean * w; the other keys' (0 - mean) * w (:478-491)
      if (normalize) bias -= w * (double)p[1 + 2 * n + total + start + j];
    }
    M.offs.push_back((int64_t)M.keys.size());
  }
  M.bias.push_back(bias);
  SortColumns(M);
  cfb_model *model = Upload(M, p, tag);
  ArmNoise(model, noise, p, rows);
  Check(cfb_predict_host(model, cols.num.data(), cols.num_sel.data(), cols.cat.data(), cols.cat_sel.data(), rows, CFB_PREDICT_SCORE,
                         FlatVector::GetData<float>(result)));
}

namespace {
// The head shared by the nb / qda parameter lists (naive_bayes.cpp:186-210, qda.cpp:367-390):
// [K | S | idx_0..idx_{S-1} | unique keys], S = m + 1 index entries when there are categorical columns.
struct ClassHead {
  size_t K = 0, S = 0, total = 0, next = 2;
  std::vector<int64_t> offs;
  std::vector<int32_t> keys;
};
ClassHead ParseHead(const std::vector<float> &p, size_t m, const char *what) {
  using namespace duckdb;
  ClassHead h;
  if (p.size() < 2) throw InvalidInputException(std::string(what) + ": parameter list too short");
  h.K = (size_t)p[0];
  h.S = (size_t)p[1];
  if (h.K < 1) throw InvalidInputException(std::string(what) + ": no classes in the parameter list");
  if ((h.S == 0) != (m == 0) || (h.S && h.S != m + 1))
    throw InvalidInputException(std::string(what) + ": the parameter list was trained on a different number of categorical columns");
  h.offs.push_back(0);
  if (h.S) {
    if (p.size() < 2 + h.S) throw InvalidInputException(std::string(what) + ": parameter list too short");
    for (size_t i = 0; i < h.S; i++) {
      const int64_t v = (int64_t)p[2 + i];
      if (i == 0 ? v != 0 : v < h.offs.back()) throw InvalidInputException(std::string(what) + ": malformed categorical index");
      if (i) h.offs.push_back(v);
    }
    h.total = (size_t)h.offs.back();
    if (p.size() < 2 + h.S + h.total) throw InvalidInputException(std::string(what) + ": parameter list too short");
    for (size_t t = 0; t < h.total; t++) h.keys.push_back((int32_t)(int64_t)p[2 + h.S + t]);
    h.next = 2 + h.S + h.total;
  }
  return h;
}
// ascending keys per column (the device's dense key maps want them sorted); perm[t] = position in the list's order
std::vector<size_t> SortKeys(ClassHead &h, size_t m) {
  std::vector<size_t> perm(h.total);
  for (size_t i = 0; i < h.total; i++) perm[i] = i;
  for (size_t c = 0; c < m; c++)
    std::sort(perm.begin() + h.offs[c], perm.begin() + h.offs[c + 1], [&](size_t x, size_t y) { return h.keys[x] < h.keys[y]; });
  std::vector<int32_t> k2(h.total);
  for (size_t i = 0; i < h.total; i++) k2[i] = h.keys[perm[i]];
  h.keys.swap(k2);
  return perm;
}
}  // namespace

// naive_bayes.cpp:153-263.  params = [K | S | idx | keys | labels[K] | priors[K] | per class: (mean, variance) per numeric
//                                     column, then one probability per (column, key)]; the result is the class LABEL (:252)
void nb_impute(duckdb::DataChunk &args, duckdb::ExpressionState &, duckdb::Vector &result) {
  using namespace duckdb;
  const idx_t rows = args.size();
  if (rows == 0) return;
  if (args.ColumnCount() < 2) throw InvalidInputException("nb_predict(params, normalize, columns...)");
  const std::vector<float> p = Params(args.data[0], rows);
  Columns cols;
  Classify(args, 2, cols);
  const size_t n = cols.num.size(), m = cols.cat.size();
  const int tag = (int)(8 | (n << 8) | (m << 16));
  cfb_model *model = Cached(p, tag);
  if (!model) {
    ClassHead h = ParseHead(p, m, "nb_predict");
    const size_t K = h.K, total = h.total, per_class = 2 * n + total, at = h.next + 2 * K;
    if (p.size() < at + K * per_class) throw InvalidInputException("nb_predict: parameter list too short for these columns");
    const std::vector<size_t> perm = SortKeys(h, m);
    std::vector<int32_t> labels(K);
    std::vector<double> prior(K), mean(K * n), var(K * n), prob(K * total);
    for (size_t k = 0; k < K; k++) {
      labels[k] = (int32_t)p[h.next + k];
      prior[k] = p[h.next + K + k];
      for (size_t j = 0; j < n; j++) {
        mean[k * n + j] = p[at + k * per_class + 2 * j];
        var[k * n + j] = p[at + k * per_class + 2 * j + 1];
      }
      for (size_t t = 0; t < total; t++) prob[k * total + t] = p[at + k * per_class + 2 * n + perm[t]];
    }
    cfb_nb_model M{};
    M.n_num = (int)n;
    M.n_cat = (int)m;
    M.n_classes = (int)K;
    M.labels = labels.data();
    M.prior = prior.data();
    M.mean = mean.data();
    M.var = var.data();
    M.cat_offsets = h.offs.data();
    M.cat_keys = h.keys.data();
    M.cat_prob = prob.data();
    cfb_model_destroy(t_cache.model);
    t_cache.model = nullptr;
    Check(cfb_model_create_nb(Device(), &M, &t_cache.model));
    t_cache.params = p;
    t_cache.tag = tag;
    model = t_cache.model;
  }
  Check(cfb_predict_host(model, cols.num.data(), cols.num_sel.data(), cols.cat.data(), cols.cat_sel.data(), rows, CFB_PREDICT_LABEL,
                         FlatVector::GetData<int32_t>(result)));
}

// qda.cpp:338-498.  params = [K | S | idx | keys | labels[K] | per class: Q[p x p] column-major, l[p], intercept |
//                             (means[p] when normalize)], p = n + number of keys; the result is the class LABEL (:480)
void qda_impute(duckdb::DataChunk &args, duckdb::ExpressionState &, duckdb::Vector &result) {
  using namespace duckdb;
  const idx_t rows = args.size();
  if (rows == 0) return;
  if (args.ColumnCount() < 2) throw InvalidInputException("qda_predict(params, normalize, columns...)");
  const std::vector<float> p = Params(args.data[0], rows);
  const bool normalize = Flag(args.data[1], rows);
  Columns cols;
  Classify(args, 2, cols);
  const size_t n = cols.num.size(), m = cols.cat.size();
  const int tag = (int)(16 | (normalize ? 2 : 0) | (n << 8) | (m << 16));
  cfb_model *model = Cached(p, tag);
  if (!model) {
    ClassHead h = ParseHead(p, m, "qda_predict");
    const size_t K = h.K, total = h.total, P = n + total, per_class = P * P + P + 1, at = h.next + K;
    if (P == 0 || p.size() < at + K * per_class + (normalize ? P : 0)) throw InvalidInputException("qda_predict: parameter list too short for these columns");
    const std::vector<size_t> perm = SortKeys(h, m);
    // feature f -> its position in the list's order (numeric columns stay, keys follow the sort)
    std::vector<size_t> src(P);
    for (size_t i = 0; i < n; i++) src[i] = i;
    for (size_t t = 0; t < total; t++) src[n + t] = n + perm[t];
    std::vector<int32_t> labels(K);
    std::vector<double> quad(K * P * P), lin(K * P), icpt(K), center(P);
    for (size_t k = 0; k < K; k++) {
      labels[k] = (int32_t)p[h.next + k];
      const size_t base = at + k * per_class;
      for (size_t j = 0; j < P; j++) {
        for (size_t i = 0; i < P; i++) quad[k * P * P + i + j * P] = p[base + src[i] + src[j] * P];
        lin[k * P + j] = p[base + P * P + src[j]];
      }
      icpt[k] = p[base + P * P + P];
    }
    if (normalize)
      for (size_t j = 0; j < P; j++) center[j] = p[at + K * per_class + src[j]];
    cfb_qda_model M{};
    M.n_num = (int)n;
    M.n_cat = (int)m;
    M.n_classes = (int)K;
    M.labels = labels.data();
    M.quad = quad.data();
    M.lin = lin.data();
    M.intercept = icpt.data();
    M.center = normalize ? center.data() : nullptr;
    M.cat_offsets = h.offs.data();
    M.cat_keys = h.keys.data();
    cfb_model_destroy(t_cache.model);
    t_cache.model = nullptr;
    Check(cfb_model_create_qda(Device(), &M, &t_cache.model));
    t_cache.params = p;
    t_cache.tag = tag;
    model = t_cache.model;
  }
  Check(cfb_predict_host(model, cols.num.data(), cols.num_sel.data(), cols.cat.data(), cols.cat_sel.data(), rows, CFB_PREDICT_LABEL,
                         FlatVector::GetData<int32_t>(result)));
}

// lda.cpp:421-590.  params = [K | S | idx_0..idx_{S-1} | unique keys | target labels[K] | coefficients[K][n + total] |
//                             intercept[K] | (means[n + total] when normalize)];  the result is the class INDEX (:575)
void LDA_impute(duckdb::DataChunk &args, duckdb::ExpressionState &, duckdb::Vector &result) {
  using namespace duckdb;
  const idx_t rows = args.size();
  if (rows == 0) return;
  if (args.ColumnCount() < 2) throw InvalidInputException("lda_predict(params, normalize, columns...)");
  const std::vector<float> p = Params(args.data[0], rows);
  const bool normalize = Flag(args.data[1], rows);
  Columns cols;
  Classify(args, 2, cols);
  const size_t n = cols.num.size(), m = cols.cat.size();
  const int tag = (int)(1 | (normalize ? 2 : 0) | (n << 8) | (m << 16));
  if (cfb_model *hit = Cached(p, tag)) {
    Check(cfb_predict_host(hit, cols.num.data(), cols.num_sel.data(), cols.cat.data(), cols.cat_sel.data(), rows,
                           CFB_PREDICT_ARGMAX, FlatVector::GetData<int32_t>(result)));
    return;
  }
  if (p.size() < 2) throw InvalidInputException("parameter list too short");
  const size_t K = (size_t)p[0], S = (size_t)p[1];
  if (K < 1 || K > 32) throw InvalidInputException("lda_predict: 1..32 classes");
  if ((S == 0) != (m == 0) || (S && S != m + 1)) throw InvalidInputException("the parameter list was trained on a different number of categorical columns");
  size_t off = 2, total = 0;
  std::vector<size_t> idx(S);
  for (size_t i = 0; i < S; i++) {
    if (off + i >= p.size()) throw InvalidInputException("parameter list too short");
    idx[i] = (size_t)p[off + i];
  }
  off += S;
  if (S) total = idx[S - 1];
  const size_t keys_at = off;
  off += total;
  off += K;  // target labels (:475-480): not used, the function returns the class index
  const size_t np = n + total, coef_at = off, icpt_at = coef_at + np * K, mean_at = icpt_at + K;
  if (p.size() < mean_at + (normalize ? np : 0)) throw InvalidInputException("parameter list too short for these columns");
  Model M;
  M.m.n_num = (int)n;
  M.m.n_cat = (int)m;
  M.m.n_out = (int)K;
  M.w_num.resize(K * n);
  M.w_cat.resize(K * total);
  M.bias.resize(K);
  for (size_t k = 0; k < K; k++) {
    double bias = p[icpt_at + k];
    for (size_t j = 0; j < np; j++) {
      const double w = p[coef_at + k * np + j];  // coefficients[(j * K) + k] = params[(k * np) + j + off]  (:487-491)
      if (j < n) M.w_num[k * n + j] = w;
      else M.w_cat[k * total + (j - n)] = w;
      if (normalize) bias -= w * (double)p[mean_at + j];  // centred features (:533-549)
    }
    M.bias[k] = bias;
  }
  M.offs.push_back(0);
  for (size_t c = 0; c < m; c++) {
    if (idx[c + 1] < idx[c] || idx[c + 1] > total) throw InvalidInputException("malformed categorical index in the parameter list");
    for (size_t j = idx[c]; j < idx[c + 1]; j++) M.keys.push_back((int32_t)p[keys_at + j]);
    M.offs.push_back((int64_t)M.keys.size());
  }
  if (m && idx[0] != 0) throw InvalidInputException("malformed categorical index in the parameter list");
  SortColumns(M);
  Check(cfb_predict_host(Upload(M, p, tag), cols.num.data(), cols.num_sel.data(), cols.cat.data(), cols.cat_sel.data(), rows,
                         CFB_PREDICT_ARGMAX, FlatVector::GetData<int32_t>(result)));
}

duckdb::unique_ptr<duckdb::FunctionData> linreg_impute_bind(duckdb::ClientContext &, duckdb::ScalarFunction &function,
                                                            duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &) {
  function.return_type = duckdb::LogicalType::FLOAT;  // regression.cpp:366-375
  function.varargs = duckdb::LogicalType::ANY;
  return duckdb::make_uniq<duckdb::VariableReturnBindData>(function.return_type);
}
duckdb::unique_ptr<duckdb::FunctionData> LDA_impute_bind(duckdb::ClientContext &, duckdb::ScalarFunction &function,
                                                         duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &) {
  function.return_type = duckdb::LogicalType::INTEGER;  // lda.cpp:593-601
  function.varargs = duckdb::LogicalType::ANY;
  return duckdb::make_uniq<duckdb::VariableReturnBindData>(function.return_type);
}
duckdb::unique_ptr<duckdb::FunctionData> nb_impute_bind(duckdb::ClientContext &, duckdb::ScalarFunction &function,
                                                        duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &) {
  function.return_type = duckdb::LogicalType::INTEGER;  // naive_bayes.cpp:265-275
  function.varargs = duckdb::LogicalType::ANY;
  return duckdb::make_uniq<duckdb::VariableReturnBindData>(function.return_type);
}
duckdb::unique_ptr<duckdb::FunctionData> qda_impute_bind(duckdb::ClientContext &, duckdb::ScalarFunction &function,
                                                         duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &) {
  function.return_type = duckdb::LogicalType::INTEGER;  // qda.cpp:500-509
  function.varargs = duckdb::LogicalType::ANY;
  return duckdb::make_uniq<duckdb::VariableReturnBindData>(function.return_type);
}

}  // namespace ML
