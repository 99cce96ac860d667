// predict_glue.cpp -- linreg_predict / lda_predict, B200 build: the write-back step of a MICE iteration.
//
//   linreg_predict(params FLOAT[], noise BOOL, normalize BOOL, cols...)  ML::linreg_impute  reference: ML/regression.cpp:397-509
//   lda_predict(params FLOAT[], normalize BOOL, cols...)                 LDA_impute         reference: ML/lda.cpp:421-590
// The parameter lists keep the reference's layout (they are what linreg_train / lda_train emit); the glue
// turns them into a cfb_linear_model -- bias, numeric weights, per-(column,key) weights, with the
// `normalize` centering folded into the bias -- and scores the chunk on the GPU (cfb_predict_host).
//
// Deliberate differences: a key the model does not know contributes 0 (the reference reads
// past the column's weights, regression.cpp:471-498, or asserts, lda.cpp:528); `noise = true` is refused:
// the reference draws from libc random() seeded from /dev/urandom (regression.cpp:376-393, :495-505), which
// no test can pin -- a counter-based device generator is the planned replacement.
#include <cmath>
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
  return t_cache.model;
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
  if (Flag(args.data[1], rows)) throw InvalidInputException("linreg_predict: noise = true is not supported by the B200 build (see DESIGN.md)");
  const bool normalize = Flag(args.data[2], rows);
  Columns cols;
  Classify(args, 3, cols);
  const size_t n = cols.num.size(), m = cols.cat.size();
  const int tag = (int)(0 | (normalize ? 2 : 0) | (n << 8) | (m << 16));
  if (cfb_model *hit = Cached(p, tag)) {
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
  Check(cfb_predict_host(Upload(M, p, tag), cols.num.data(), cols.num_sel.data(), cols.cat.data(), cols.cat_sel.data(), rows,
                         CFB_PREDICT_SCORE, FlatVector::GetData<float>(result)));
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

}  // namespace ML
