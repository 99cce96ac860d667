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
#include "triple_view.h"

namespace Triple {

namespace {

// One row of a triple STRUCT vector as a cfb_result that owns its arrays.
struct OwnedResult {
  cfb_result r;
  std::vector<double> lin, quad, numcat;
  std::vector<int64_t> cat_off, cat_cnt, pair_off, pair_cnt;
  std::vector<int32_t> cat_key, k1, k2;
  void Bind() {
    r.lin = lin.data();
    r.quad = quad.data();
    r.cat_offsets = cat_off.data();
    r.cat_keys = cat_key.data();
    r.cat_counts = cat_cnt.data();
    r.numcat_sums = numcat.data();
    r.pair_offsets = pair_off.data();
    r.pair_key1 = k1.data();
    r.pair_key2 = k2.data();
    r.pair_counts = pair_cnt.data();
  }
};

// One row of a triple STRUCT argument -> OwnedResult.  After a join the argument is rarely a plain flat
// vector (a CROSS JOIN hands one side over as a CONSTANT vector; a hash join slices the STRUCT's children into
// DICTIONARY vectors); the reference flattens both arguments first (mul.cpp:24-28), TripleView reads them in place.
struct TripleReader {
  TripleView v;
  TripleReader(duckdb::Vector &vec, idx_t count, bool nb) : v(vec, count, nb) {}

  void Row(idx_t row, OwnedResult &o) const {
    using namespace duckdb;
    const bool nb = v.nb;
    memset(&o.r, 0, sizeof(o.r));
    const idx_t sr = v.Row(row);
    const list_entry_t le = v.lin.Entry(sr), qe = v.quad.Entry(sr), lco = v.lin_cat.Outer(sr);
    const idx_t n = le.length, m = lco.length;
    const idx_t nq = nb ? n : n * (n + 1) / 2;
    if (n > CFB_MAX_NUM || m > CFB_MAX_CAT) throw InvalidInputException("ring product: too many columns");
    if (qe.length != nq) throw InvalidInputException("triple STRUCT lists have the wrong length");
    o.r.kind = nb ? CFB_NB : CFB_TRIPLE;
    o.r.n_num = (int)n;
    o.r.n_cat = (int)m;
    o.r.N = v.N.At<int32_t>(sr);
    o.r.n_quad = (int64_t)nq;
    o.lin.resize(n);
    o.quad.resize(nq);
    for (idx_t k = 0; k < n; k++) o.lin[k] = v.lin.elems.At<float>(le.offset + k);
    for (idx_t k = 0; k < nq; k++) o.quad[k] = v.quad.elems.At<float>(qe.offset + k);
    o.cat_off.assign(m + 1, 0);
    o.cat_key.clear();
    o.cat_cnt.clear();
    for (idx_t c = 0; c < m; c++) {
      const list_entry_t e = v.lin_cat.Inner(lco.offset + c);
      for (idx_t t = 0; t < e.length; t++) {
        o.cat_key.push_back(v.lin_cat.Leaf<int32_t>(0, e.offset + t));
        o.cat_cnt.push_back((int64_t)v.lin_cat.Leaf<float>(1, e.offset + t));  // counts are integral floats
      }
      o.cat_off[c + 1] = (int64_t)o.cat_key.size();
    }
    const idx_t tk = o.cat_key.size();
    o.r.total_keys = (int64_t)tk;
    o.numcat.clear();
    o.pair_off.assign(1, 0);
    o.k1.clear();
    o.k2.clear();
    o.pair_cnt.clear();
    if (!nb) {
      const list_entry_t nco = v.num_cat.Outer(sr), cco = v.cat_cat.Outer(sr);
      if (nco.length != n * m || cco.length != m * (m + 1) / 2)
        throw InvalidInputException("triple STRUCT lists have the wrong length");
      o.numcat.assign(n * tk, 0.0);
      for (idx_t i = 0; i < n; i++)
        for (idx_t c = 0; c < m; c++) {  // sub-list num*m + cat, same keys / order as lin_cat[cat]
          const list_entry_t e = v.num_cat.Inner(nco.offset + i * m + c);
          if ((int64_t)e.length != o.cat_off[c + 1] - o.cat_off[c])
            throw InvalidInputException("quad_num_cat and lin_cat disagree on the keys of a column");
          for (idx_t t = 0; t < e.length; t++) o.numcat[i * tk + o.cat_off[c] + t] = v.num_cat.Leaf<float>(1, e.offset + t);
        }
      const idx_t npl = m * (m + 1) / 2;
      o.r.n_pair_lists = (int64_t)npl;
      for (idx_t p = 0; p < npl; p++) {
        const list_entry_t e = v.cat_cat.Inner(cco.offset + p);
        for (idx_t t = 0; t < e.length; t++) {
          o.k1.push_back(v.cat_cat.Leaf<int32_t>(0, e.offset + t));
          o.k2.push_back(v.cat_cat.Leaf<int32_t>(1, e.offset + t));
          o.pair_cnt.push_back((int64_t)v.cat_cat.Leaf<float>(2, e.offset + t));
        }
        o.pair_off.push_back((int64_t)o.k1.size());
      }
    }
    o.Bind();
  }
};

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
