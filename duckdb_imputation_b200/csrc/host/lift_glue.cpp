// lift_glue.cpp -- the scalar lifts to_cofactor / to_nb_agg and the aggregates over lifted
// triples sum_triple / sum_nb_agg, B200 build.
//
//   to_cofactor(cols...)   Triple::CustomLift      reference: lift.cpp:15-243, bind :246-284
//   to_nb_agg(cols...)     Triple::to_nb_lift      reference: lift_to_nb_agg.cpp:13-136
//       Host scalar functions: 1 row -> one STRUCT of O(n^2) values written into DuckDB vectors.
//       Pure expansion into host memory -- there is nothing for the GPU to win.
//   sum_triple(STRUCT)     Triple::Sum             reference: sum.cpp:57-261, bind :15-53
//   sum_nb_agg(STRUCT)     Triple::sum_nb_agg      reference: sum_nb_agg.cpp:45-175
//       update = hand the flattened STRUCT children to cfb_ctx_append_triples (device column sums
//       + device scatter-add of the sparse entries); combine / finalize are the shared
//       Triple::SumStateCombine / SumStateFinalize.
#include <cstring>
#include <mutex>
#include <string>

#include "../../../include/cofactor_b200.h"
#include "triple_glue.h"
#include "triple_view.h"

namespace Triple {

namespace {

void CheckRc(int rc) {
  if (rc == CFB_OK) return;
  const std::string msg = cfb_last_error();
  if (rc == CFB_ERR_INVALID || rc == CFB_ERR_DOMAIN) throw duckdb::InvalidInputException(msg);
  throw duckdb::InternalException(msg);
}

duckdb::LogicalType KV() {
  duckdb::child_list_t<duckdb::LogicalType> kv;
  kv.emplace_back("key", duckdb::LogicalType::INTEGER);
  kv.emplace_back("value", duckdb::LogicalType::FLOAT);
  return duckdb::LogicalType::LIST(duckdb::LogicalType::LIST(duckdb::LogicalType::STRUCT(kv)));
}

// lift.cpp:255-257 names the numeric fields lin_num / quad_num; the sums name them lin_agg /
// quad_agg (sum.cpp:28-29).  sum_triple reads children positionally, so both are accepted.
duckdb::LogicalType TripleType(bool nb, const char *lin, const char *quad) {
  duckdb::child_list_t<duckdb::LogicalType> f;
  f.emplace_back("N", duckdb::LogicalType::INTEGER);
  f.emplace_back(lin, duckdb::LogicalType::LIST(duckdb::LogicalType::FLOAT));
  f.emplace_back(quad, duckdb::LogicalType::LIST(duckdb::LogicalType::FLOAT));
  f.emplace_back("lin_cat", KV());
  if (!nb) {
    f.emplace_back("quad_num_cat", KV());
    duckdb::child_list_t<duckdb::LogicalType> kkv;
    kkv.emplace_back("key1", duckdb::LogicalType::INTEGER);
    kkv.emplace_back("key2", duckdb::LogicalType::INTEGER);
    kkv.emplace_back("value", duckdb::LogicalType::FLOAT);
    f.emplace_back("quad_cat", duckdb::LogicalType::LIST(duckdb::LogicalType::LIST(duckdb::LogicalType::STRUCT(kkv))));
  }
  return duckdb::LogicalType::STRUCT(f);
}

struct KeyValueOut {
  duckdb::list_entry_t *outer, *inner;
  int32_t *key;
  float *val;
};
// Size LIST(LIST(STRUCT(key,value))) for `outer_n` outer entries / `inner_n` singleton lists.
KeyValueOut PrepareKV(duckdb::Vector &v, idx_t rows, idx_t lists_per_row) {
  using namespace duckdb;
  const idx_t n_lists = rows * lists_per_row;
  ListVector::Reserve(v, n_lists);
  ListVector::SetListSize(v, n_lists);
  Vector &inner = ListVector::GetEntry(v);
  ListVector::Reserve(inner, n_lists);  // every inner list of a lifted row is a singleton
  ListVector::SetListSize(inner, n_lists);
  auto &kv = StructVector::GetEntries(ListVector::GetEntry(inner));
  return {ListVector::GetData(v), ListVector::GetData(inner), FlatVector::GetData<int32_t>(*kv[0]),
          FlatVector::GetData<float>(*kv[1])};
}

void Lift(bool nb, duckdb::DataChunk &args, duckdb::Vector &result) {
  using namespace duckdb;
  const idx_t rows = args.size();
  const idx_t cols = args.ColumnCount();
  std::vector<UnifiedVectorFormat> fmt(cols);
  std::vector<const float *> num;
  std::vector<const int32_t *> cat;
  std::vector<const SelectionVector *> num_sel, cat_sel;
  for (idx_t j = 0; j < cols; j++) {
    args.data[j].ToUnifiedFormat(rows, fmt[j]);
    const auto &t = args.data[j].GetType();
    if (t == LogicalType::FLOAT || t == LogicalType::DOUBLE) {
      num.push_back(UnifiedVectorFormat::GetData<float>(fmt[j]));
      num_sel.push_back(fmt[j].sel);
    } else {
      cat.push_back(UnifiedVectorFormat::GetData<int32_t>(fmt[j]));
      cat_sel.push_back(fmt[j].sel);
    }
  }
  const idx_t n = num.size(), m = cat.size();
  const idx_t nq = nb ? n : n * (n + 1) / 2;
  auto &kids = StructVector::GetEntries(result);
  auto N = FlatVector::GetData<int32_t>(*kids[0]);
  ListVector::Reserve(*kids[1], n * rows);
  ListVector::SetListSize(*kids[1], n * rows);
  ListVector::Reserve(*kids[2], nq * rows);
  ListVector::SetListSize(*kids[2], nq * rows);
  auto lin_e = ListVector::GetData(*kids[1]);
  auto quad_e = ListVector::GetData(*kids[2]);
  auto lin_d = FlatVector::GetData<float>(ListVector::GetEntry(*kids[1]));
  auto quad_d = FlatVector::GetData<float>(ListVector::GetEntry(*kids[2]));
  KeyValueOut lc = PrepareKV(*kids[3], rows, m);
  KeyValueOut nc{};
  list_entry_t *cc_outer = nullptr, *cc_inner = nullptr;
  int32_t *cc_k1 = nullptr, *cc_k2 = nullptr;
  float *cc_v = nullptr;
  const idx_t npl = m * (m + 1) / 2;
  if (!nb) {
    nc = PrepareKV(*kids[4], rows, n * m);
    Vector &v = *kids[5];
    ListVector::Reserve(v, rows * npl);
    ListVector::SetListSize(v, rows * npl);
    Vector &inner = ListVector::GetEntry(v);
    ListVector::Reserve(inner, rows * npl);
    ListVector::SetListSize(inner, rows * npl);
    auto &kkv = StructVector::GetEntries(ListVector::GetEntry(inner));
    cc_outer = ListVector::GetData(v);
    cc_inner = ListVector::GetData(inner);
    cc_k1 = FlatVector::GetData<int32_t>(*kkv[0]);
    cc_k2 = FlatVector::GetData<int32_t>(*kkv[1]);
    cc_v = FlatVector::GetData<float>(*kkv[2]);
  }
  float x[CFB_MAX_NUM];
  int32_t key[CFB_MAX_CAT];
  if (n > CFB_MAX_NUM || m > CFB_MAX_CAT) throw InvalidInputException("too many columns for a ring lift");
  for (idx_t r = 0; r < rows; r++) {
    for (idx_t k = 0; k < n; k++) x[k] = num[k][num_sel[k]->get_index(r)];
    for (idx_t k = 0; k < m; k++) key[k] = cat[k][cat_sel[k]->get_index(r)];
    N[r] = 1;
    lin_e[r] = {r * n, n};
    for (idx_t k = 0; k < n; k++) lin_d[r * n + k] = x[k];
    quad_e[r] = {r * nq, nq};
    if (nb) {
      for (idx_t k = 0; k < n; k++) quad_d[r * nq + k] = x[k] * x[k];  // lift_to_nb_agg.cpp:108-117
    } else {
      idx_t p = r * nq;
      for (idx_t i = 0; i < n; i++)
        for (idx_t j = i; j < n; j++) quad_d[p++] = x[i] * x[j];  // lift.cpp:119-136
    }
    lc.outer[r] = {r * m, m};
    for (idx_t k = 0; k < m; k++) {  // lin_cat[k] = [{key, 1}]  (lift.cpp:94-104)
      const idx_t e = r * m + k;
      lc.inner[e] = {e, 1};
      lc.key[e] = key[k];
      lc.val[e] = 1.f;
    }
    if (nb) continue;
    nc.outer[r] = {r * n * m, n * m};
    for (idx_t i = 0; i < n; i++)
      for (idx_t k = 0; k < m; k++) {  // quad_num_cat[i*m+k] = [{key_k, x_i}]  (lift.cpp:157-176)
        const idx_t e = (r * n + i) * m + k;
        nc.inner[e] = {e, 1};
        nc.key[e] = key[k];
        nc.val[e] = x[i];
      }
    cc_outer[r] = {r * npl, npl};
    idx_t e = r * npl;
    for (idx_t k = 0; k < m; k++)
      for (idx_t l = k; l < m; l++, e++) {  // quad_cat[(k<=l)] = [{key_k, key_l, 1}]  (lift.cpp:199-219)
        cc_inner[e] = {e, 1};
        cc_k1[e] = key[k];
        cc_k2[e] = key[l];
        cc_v[e] = 1.f;
      }
  }
}

// update of sum_triple / sum_nb_agg: route rows to states, hand each state its rows' children.  The STRUCT
// argument may be FLAT, CONSTANT or DICTIONARY at any nesting level (a join or a filter below the aggregate):
// the reference flattens it first (sum.cpp:72 -> utils.cpp:3-18); TripleView reads it in place.
void SumLifted(int kind, duckdb::Vector inputs[], idx_t input_count, duckdb::Vector &state_vector, idx_t count) {
  using namespace duckdb;
  if (input_count != 1) throw InvalidInputException("sum over lifted triples takes one STRUCT argument");
  if (count == 0) return;
  UnifiedVectorFormat sdata;
  state_vector.ToUnifiedFormat(count, sdata);
  auto states = (SumState **)sdata.data;
  const bool nb = kind == CFB_NB;
  const TripleView in(inputs[0], count, nb);
  // shape from the first row (sum.cpp:96-106): n = |lin|, m = |lin_cat|
  const idx_t n = in.lin.Entry(in.Row(0)).length, m = in.lin_cat.Outer(in.Row(0)).length;
  const idx_t nq = nb ? n : n * (n + 1) / 2, npl = m * (m + 1) / 2;
  if (n > CFB_MAX_NUM || m > CFB_MAX_CAT) throw InvalidInputException("too many columns in a lifted triple");
  // the sparse leaves as flat arrays (the vectors' own buffers unless something below was sliced)
  std::vector<int32_t> s_lck, s_nck, s_cc1, s_cc2;
  std::vector<float> s_lcv, s_ncv, s_ccv;
  const int32_t *lc_key = in.lin_cat.FlatLeaf<int32_t>(0, s_lck);
  const float *lc_val = in.lin_cat.FlatLeaf<float>(1, s_lcv);
  const int32_t *nc_key = nullptr, *cc_k1 = nullptr, *cc_k2 = nullptr;
  const float *nc_val = nullptr, *cc_val = nullptr;
  if (!nb) {
    nc_key = in.num_cat.FlatLeaf<int32_t>(0, s_nck);
    nc_val = in.num_cat.FlatLeaf<float>(1, s_ncv);
    cc_k1 = in.cat_cat.FlatLeaf<int32_t>(0, s_cc1);
    cc_k2 = in.cat_cat.FlatLeaf<int32_t>(1, s_cc2);
    cc_val = in.cat_cat.FlatLeaf<float>(2, s_ccv);
  }
  // Gather per state: compact copies of the numeric children and the inner list entries of the
  // rows that belong to it (a chunk normally has one state, or a handful with GROUP BY).
  std::vector<SumState *> order;
  std::vector<std::vector<idx_t>> rows_of;
  for (idx_t r = 0; r < count; r++) {
    SumState *s = states[sdata.sel->get_index(r)];
    size_t b = 0;
    while (b < order.size() && order[b] != s) b++;
    if (b == order.size()) {
      order.push_back(s);
      rows_of.emplace_back();
    }
    rows_of[b].push_back(r);
  }
  std::vector<int32_t> gN;
  std::vector<float> gl, gq;
  std::vector<cfb_list_entry> glc, gnc, gcc;
  for (size_t b = 0; b < order.size(); b++) {
    const auto &rows = rows_of[b];
    gN.clear(); gl.clear(); gq.clear(); glc.clear(); gnc.clear(); gcc.clear();
    for (idx_t r : rows) {
      const idx_t sr = in.Row(r);
      const list_entry_t le = in.lin.Entry(sr), qe = in.quad.Entry(sr), lco = in.lin_cat.Outer(sr);
      if (le.length != n || qe.length != nq || lco.length != m)
        throw InvalidInputException("triples of different shapes in one aggregate");
      gN.push_back(in.N.At<int32_t>(sr));
      for (idx_t k = 0; k < n; k++) gl.push_back(in.lin.elems.At<float>(le.offset + k));
      for (idx_t k = 0; k < nq; k++) gq.push_back(in.quad.elems.At<float>(qe.offset + k));
      for (idx_t k = 0; k < m; k++) {
        const list_entry_t e = in.lin_cat.Inner(lco.offset + k);
        glc.push_back({e.offset, e.length});
      }
      if (nb) continue;
      const list_entry_t nco = in.num_cat.Outer(sr), cco = in.cat_cat.Outer(sr);
      if (nco.length != n * m || cco.length != npl) throw InvalidInputException("triple STRUCT lists have the wrong length");
      for (idx_t e = 0; e < n * m; e++) {
        const list_entry_t x = in.num_cat.Inner(nco.offset + e);
        gnc.push_back({x.offset, x.length});
      }
      for (idx_t e = 0; e < npl; e++) {
        const list_entry_t x = in.cat_cat.Inner(cco.offset + e);
        gcc.push_back({x.offset, x.length});
      }
    }
    // a state that is fed alone keeps a private context; GROUP BY states share their worker thread's arena
    SumState *s = order[b];
    if (!s->arena) {
      if (order.size() == 1) PrivateContext(*s, kind, (int)n, (int)m);
      else AssignSlot(*s, kind, (int)n, (int)m);
    }
    if (s->arena->kind != kind || s->arena->n != (int)n || s->arena->m != (int)m)
      throw InvalidInputException("triples of different shapes in one aggregate");
    std::lock_guard<std::mutex> g(s->arena->mu);
    CheckRc(cfb_ctx_append_triples_slot(s->arena->ctx, s->slot, rows.size(), gN.data(), gl.data(), gq.data(), glc.data(), lc_key,
                                        lc_val, nb ? nullptr : gnc.data(), nc_key, nc_val, nb ? nullptr : gcc.data(), cc_k1,
                                        cc_k2, cc_val));
  }
}

// simple_update of sum_triple / sum_nb_agg: every row of the chunk belongs to the one state
void SumLiftedSimple(int kind, duckdb::Vector inputs[], idx_t input_count, duckdb::data_ptr_t state, idx_t count) {
  duckdb::Vector states(duckdb::LogicalType::POINTER, (duckdb::data_ptr_t)&state);
  states.SetVectorType(duckdb::VectorType::CONSTANT_VECTOR);
  SumLifted(kind, inputs, input_count, states, count);
}

}  // namespace

void CustomLift(duckdb::DataChunk &args, duckdb::ExpressionState &, duckdb::Vector &result) { Lift(false, args, result); }
void to_nb_lift(duckdb::DataChunk &args, duckdb::ExpressionState &, duckdb::Vector &result) { Lift(true, args, result); }

duckdb::unique_ptr<duckdb::FunctionData> CustomLiftBind(duckdb::ClientContext &, duckdb::ScalarFunction &function,
                                                        duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &) {
  function.return_type = TripleType(false, "lin_num", "quad_num");
  function.varargs = duckdb::LogicalType::ANY;
  return duckdb::make_uniq<duckdb::VariableReturnBindData>(function.return_type);
}
duckdb::unique_ptr<duckdb::FunctionData> to_nb_lift_bind(duckdb::ClientContext &, duckdb::ScalarFunction &function,
                                                         duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &) {
  function.return_type = TripleType(true, "lin_num", "quad_num");
  function.varargs = duckdb::LogicalType::ANY;
  return duckdb::make_uniq<duckdb::VariableReturnBindData>(function.return_type);
}

duckdb::unique_ptr<duckdb::FunctionData> SumBind(duckdb::ClientContext &, duckdb::AggregateFunction &function,
                                                 duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &) {
  function.return_type = TripleType(false, "lin_agg", "quad_agg");
  return duckdb::make_uniq<duckdb::VariableReturnBindData>(function.return_type);
}
duckdb::unique_ptr<duckdb::FunctionData> sum_nb_agg_bind(duckdb::ClientContext &, duckdb::AggregateFunction &function,
                                                         duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &) {
  function.return_type = TripleType(true, "lin_agg", "quad_agg");
  return duckdb::make_uniq<duckdb::VariableReturnBindData>(function.return_type);
}

void Sum(duckdb::Vector inputs[], duckdb::AggregateInputData &, idx_t input_count, duckdb::Vector &state_vector, idx_t count) {
  SumLifted(CFB_TRIPLE, inputs, input_count, state_vector, count);
}
void sum_nb_agg(duckdb::Vector inputs[], duckdb::AggregateInputData &, idx_t input_count, duckdb::Vector &state_vector,
                idx_t count) {
  SumLifted(CFB_NB, inputs, input_count, state_vector, count);
}
void SumSimple(duckdb::Vector inputs[], duckdb::AggregateInputData &, idx_t input_count, duckdb::data_ptr_t state, idx_t count) {
  SumLiftedSimple(CFB_TRIPLE, inputs, input_count, state, count);
}
void sum_nb_agg_simple(duckdb::Vector inputs[], duckdb::AggregateInputData &, idx_t input_count, duckdb::data_ptr_t state,
                       idx_t count) {
  SumLiftedSimple(CFB_NB, inputs, input_count, state, count);
}

}  // namespace Triple
