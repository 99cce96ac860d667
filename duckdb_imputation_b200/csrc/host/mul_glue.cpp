// mul_glue.cpp -- multiply_triple / multiply_nb_agg, B200 build.
//
//   multiply_triple(A, B)   Triple::MultiplyFunction   reference: mul.cpp:17-611, bind :614-650
//   multiply_nb_agg(A, B)   Triple::multiply_nb        reference: mul_nb.cpp
// The ring product of two per-group results for factorised joins (README.md:165-173): a scalar
// function over a join's output rows, each row two small STRUCTs.  Host code on O(result) values
// per row, like the reference; the arithmetic is cfb_result_multiply of the C ABI.
// Result fields are named lin_num / quad_num (mul.cpp:621-623), as to_cofactor's.
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/cofactor_b200.h"
#include "triple_glue.h"
#include "triple_reader.h"
#include "triple_view.h"

namespace Triple {

namespace {

void Multiply(bool nb, duckdb::DataChunk &args, duckdb::Vector &result) {
  using namespace duckdb;
  if (args.ColumnCount() != 2) throw InvalidInputException("the ring product takes two triples");
  const idx_t rows = args.size();
  if (rows == 0) return;
  TripleReader A(args.data[0], rows, nb), B(args.data[1], rows, nb);
  std::vector<cfb_result> res(rows);
  struct Guard {
    std::vector<cfb_result> &r;
    ~Guard() {
      for (auto &x : r) cfb_result_free(&x);
    }
  } guard{res};
  for (auto &x : res) memset(&x, 0, sizeof(x));
  OwnedResult a, b;
  int n = 0, m = 0;
  for (idx_t r = 0; r < rows; r++) {
    A.Row(r, a);
    B.Row(r, b);
    if (cfb_result_multiply(&a.r, &b.r, &res[r]) != CFB_OK) throw InvalidInputException(cfb_last_error());
    if (r && (res[r].n_num != n || res[r].n_cat != m)) throw InvalidInputException("triples of different shapes in one column");
    n = res[r].n_num;
    m = res[r].n_cat;
  }
  WriteResults(res, result, nb, n, m);
}

duckdb::LogicalType ProductType(bool nb) {
  using namespace duckdb;
  child_list_t<LogicalType> kv;
  kv.emplace_back("key", LogicalType::INTEGER);
  kv.emplace_back("value", LogicalType::FLOAT);
  child_list_t<LogicalType> f;
  f.emplace_back("N", LogicalType::INTEGER);
  f.emplace_back("lin_num", LogicalType::LIST(LogicalType::FLOAT));
  f.emplace_back("quad_num", LogicalType::LIST(LogicalType::FLOAT));
  f.emplace_back("lin_cat", LogicalType::LIST(LogicalType::LIST(LogicalType::STRUCT(kv))));
  if (!nb) {
    f.emplace_back("quad_num_cat", LogicalType::LIST(LogicalType::LIST(LogicalType::STRUCT(kv))));
    child_list_t<LogicalType> kkv;
    kkv.emplace_back("key1", LogicalType::INTEGER);
    kkv.emplace_back("key2", LogicalType::INTEGER);
    kkv.emplace_back("value", LogicalType::FLOAT);
    f.emplace_back("quad_cat", LogicalType::LIST(LogicalType::LIST(LogicalType::STRUCT(kkv))));
  }
  return LogicalType::STRUCT(f);
}

}  // namespace

void MultiplyFunction(duckdb::DataChunk &args, duckdb::ExpressionState &, duckdb::Vector &result) { Multiply(false, args, result); }
void multiply_nb(duckdb::DataChunk &args, duckdb::ExpressionState &, duckdb::Vector &result) { Multiply(true, args, result); }

duckdb::unique_ptr<duckdb::FunctionData> MultiplyBind(duckdb::ClientContext &, duckdb::ScalarFunction &function,
                                                      duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &) {
  function.return_type = ProductType(false);
  return duckdb::make_uniq<duckdb::VariableReturnBindData>(function.return_type);
}
duckdb::unique_ptr<duckdb::FunctionData> multiply_nb_bind(duckdb::ClientContext &, duckdb::ScalarFunction &function,
                                                          duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &) {
  function.return_type = ProductType(true);
  return duckdb::make_uniq<duckdb::VariableReturnBindData>(function.return_type);
}

}  // namespace Triple
