// train_glue.cpp -- linreg_train / lda_train, B200 build.
//
//   linreg_train(triple, label, step, lambda, max_iterations, variance, normalize)
//                                    ML::ridge_linear_regression   reference: ML/regression.cpp:113-356, bind :357-363
//   lda_train(triple, label, shrinkage, normalize)
//                                    lda_train                     reference: ML/lda.cpp:154-410, bind :146-152
// Scalar functions over ONE finished triple (a constant STRUCT): the row is read through TripleView (any vector
// shape), the sigma matrix is assembled and the model is solved on the device (cfb_sigma_from_result,
// cfb_sigma_linreg_train, cfb_sigma_lda_train); the FLOAT[] parameter list keeps the reference's layout -- it is what
// linreg_predict / lda_predict read.
#include <cmath>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include "../../../include/cofactor_b200.h"
#include "triple_glue.h"
#include "triple_reader.h"

namespace {

int TrainDevice() {
  const char *one = getenv("CFB_DEVICE");
  return one ? atoi(one) : 0;
}

void Check(int rc) {
  if (rc != CFB_OK) throw duckdb::InvalidInputException(cfb_last_error());
}

struct SigmaHandle {
  cfb_sigma *s = nullptr;
  ~SigmaHandle() { cfb_sigma_destroy(s); }
};

template <class T>
T Constant(duckdb::DataChunk &args, idx_t col, const char *fn) {
  if (col >= args.ColumnCount()) throw duckdb::InvalidInputException(std::string(fn) + ": too few arguments");
  return args.data[col].GetValue(0).GetValue<T>();
}

void EmitList(duckdb::Vector &result, const std::vector<float> &d) {
  using namespace duckdb;
  result.SetVectorType(VectorType::CONSTANT_VECTOR);
  ListVector::Reserve(result, d.size());
  ListVector::SetListSize(result, d.size());
  auto out = ConstantVector::GetData<float>(ListVector::GetEntry(result));
  for (size_t i = 0; i < d.size(); i++) out[i] = d[i];
  ListVector::GetData(result)[0] = {0, d.size()};
}

}  // namespace

namespace ML {

void ridge_linear_regression(duckdb::DataChunk &args, duckdb::ExpressionState &, duckdb::Vector &result) {
  using namespace duckdb;
  const char *fn = "linreg_train";
  const int label = Constant<int>(args, 1, fn);
  const float step_size = Constant<float>(args, 2, fn), lambda = Constant<float>(args, 3, fn);
  const int max_iterations = Constant<int>(args, 4, fn);
  const bool compute_variance = Constant<bool>(args, 5, fn), normalize = Constant<bool>(args, 6, fn);
  Triple::TripleReader reader(args.data[0], std::max<idx_t>(1, args.size()), false);
  Triple::OwnedResult row;
  reader.Row(0, row);
  if (label < 0 || label >= row.r.n_num)  // the reference only prints "label ID >= number of continuous attributes" (:137)
    throw InvalidInputException("linreg_train: label " + std::to_string(label) + " is not a numeric column of the triple");
  SigmaHandle h;
  Check(cfb_sigma_from_result(TrainDevice(), &row.r, -1, 0, &h.s));
  int32_t p = 0;
  int64_t n_values = 0;
  Check(cfb_sigma_shape(h.s, &p, nullptr, &n_values));
  const int m = row.r.n_cat;
  std::vector<int64_t> cat_array((size_t)n_values);
  std::vector<int32_t> idxs((size_t)m + 1);
  Check(cfb_sigma_layout(h.s, cat_array.data(), idxs.data()));
  std::vector<double> coeff((size_t)p), means((size_t)p);
  double variance = 0.0;
  Check(cfb_sigma_linreg_train(h.s, label, step_size, lambda, max_iterations, normalize, coeff.data(), means.data(), &variance, nullptr, nullptr));
  // the parameter list (regression.cpp:276-354)
  std::vector<float> d;
  d.push_back((float)m);
  if (m > 0) {
    for (int i = 0; i <= m; i++) d.push_back((float)idxs[(size_t)i]);
    for (auto k : cat_array) d.push_back((float)(uint64_t)k);
  }
  const int lab = label + 1;
  for (int i = 0; i < p; i++)
    if (i != lab) d.push_back((float)coeff[(size_t)i]);
  if (normalize)
    for (int i = 1; i < p; i++)
      if (i != lab) d.push_back((float)means[(size_t)i]);
  if (compute_variance) d.push_back((float)std::sqrt(variance));  // "returns std instead of variance" (:350)
  EmitList(result, d);
}

duckdb::unique_ptr<duckdb::FunctionData> ridge_linear_regression_bind(duckdb::ClientContext &, duckdb::ScalarFunction &function,
                                                                      duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &) {
  function.return_type = duckdb::LogicalType::LIST(duckdb::LogicalType::FLOAT);
  return duckdb::make_uniq<duckdb::VariableReturnBindData>(function.return_type);
}

}  // namespace ML

void lda_train(duckdb::DataChunk &args, duckdb::ExpressionState &, duckdb::Vector &result) {
  using namespace duckdb;
  const char *fn = "lda_train";
  const int label = Constant<int>(args, 1, fn);
  const float shrinkage = Constant<float>(args, 2, fn);
  const bool normalize = Constant<bool>(args, 3, fn);
  Triple::TripleReader reader(args.data[0], std::max<idx_t>(1, args.size()), false);
  Triple::OwnedResult row;
  reader.Row(0, row);
  const int n = row.r.n_num, m = row.r.n_cat;
  if (label < 0 || label >= m) throw InvalidInputException("lda_train: label " + std::to_string(label) + " is not a categorical column of the triple");
  SigmaHandle h;
  Check(cfb_sigma_from_result(TrainDevice(), &row.r, label, 0, &h.s));
  int32_t p = 0, C = 0;
  int64_t n_values = 0;
  Check(cfb_sigma_shape(h.s, &p, &C, &n_values));
  std::vector<int64_t> cat_array((size_t)n_values);
  std::vector<int32_t> idxs((size_t)m + 1);
  Check(cfb_sigma_layout(h.s, cat_array.data(), idxs.data()));
  const int q = p - 1;
  std::vector<double> coef((size_t)C * q), intercept((size_t)C), means((size_t)p);
  Check(cfb_sigma_lda_train(h.s, shrinkage, normalize, coef.data(), intercept.data(), means.data()));
  // the parameter list (lda.cpp:335-385)
  std::vector<float> d;
  d.push_back((float)C);
  d.push_back((float)(m == 1 ? 0 : m));
  if (q - n > 0) {  // categorical features besides the label: their offsets without the label's keys, then their keys
    int remove = 0;
    for (int i = 0; i <= m; i++) {
      if (i == label) {
        remove = idxs[(size_t)label + 1] - idxs[(size_t)label];
        continue;
      }
      d.push_back((float)(idxs[(size_t)i] - remove));
    }
    for (int i = 0; i < idxs[(size_t)label]; i++) d.push_back((float)(uint64_t)cat_array[(size_t)i]);
    for (int i = idxs[(size_t)label + 1]; i < idxs[(size_t)m]; i++) d.push_back((float)(uint64_t)cat_array[(size_t)i]);
  }
  for (int i = idxs[(size_t)label]; i < idxs[(size_t)label + 1]; i++) d.push_back((float)(uint64_t)cat_array[(size_t)i]);
  for (double v : coef) d.push_back((float)v);
  for (double v : intercept) d.push_back((float)v);
  if (normalize)
    for (int i = 0; i < q; i++) d.push_back((float)means[(size_t)i + 1]);
  EmitList(result, d);
}

duckdb::unique_ptr<duckdb::FunctionData> lda_train_bind(duckdb::ClientContext &, duckdb::ScalarFunction &function,
                                                        duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &) {
  function.return_type = duckdb::LogicalType::LIST(duckdb::LogicalType::FLOAT);
  return duckdb::make_uniq<duckdb::VariableReturnBindData>(function.return_type);
}
