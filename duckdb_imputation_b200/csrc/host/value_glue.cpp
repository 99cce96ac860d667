// value_glue.cpp -- Triple::sum_triple / subtract_triple / sum_nb_triple over duckdb::Value (see value_glue.h).
//
// A ring STRUCT value {N, lin_num, quad_num, lin_cat, quad_num_cat, quad_cat} (children are read by position, as the
// reference does: sum.cpp:71-75) is unpacked into a cfb_result, combined with cfb_result_combine(.., +1 / -1,
// CFB_COMBINE_KEEP_ZERO_KEYS) and packed again with the field names and FLOAT / INTEGER narrowing of the reference's
// outputs (sum.cpp:79, :97, :117, :148, :176, :206).  What the reference's std::map merges do is kept:
//   * lists come back in ascending key / (key1, key2) order                              (sum.cpp:11-35, :38-62)
//   * a key that only one operand has is kept (sum) ; a key whose count becomes 0 is kept (sub.cpp:14-38)
//   * an operand whose lists are all empty is the ring's zero: the other operand's lists are taken (sum.cpp:86-93 ..)
// Where the reference is not a ring operation the mirror is: `zero - b` is -b here (the reference returns +b's lists
// with N = -N_b, sub.cpp:93-96), `a - zero` with categorical columns is a here (the reference asks a LIST for a FLOAT
// and DuckDB throws, sub.cpp:135-138), and a key that only the subtrahend has gets the negative count (the reference
// prints "Error, key is not present in first triple" and skips it, sub.cpp:25-26).
#include "value_glue.h"

#include <algorithm>
#include <map>
#include <set>

#include "../../../include/cofactor_b200.h"

using duckdb::Value;
using duckdb::vector;

namespace {

struct Unpacked {  // owns the arrays a cfb_result points into
  cfb_result r{};
  std::vector<double> lin, quad, numcat;
  std::vector<int64_t> cat_offsets, cat_counts, pair_offsets, pair_counts;
  std::vector<int32_t> cat_keys, pair_key1, pair_key2;
  bool zero = false;  // all lists empty: the ring's zero, shape unknown
  void Bind() {
    r.lin = lin.data();
    r.quad = quad.data();
    r.cat_offsets = cat_offsets.data();
    r.cat_keys = cat_keys.data();
    r.cat_counts = cat_counts.data();
    r.numcat_sums = numcat.data();
    r.pair_offsets = pair_offsets.data();
    r.pair_key1 = pair_key1.data();
    r.pair_key2 = pair_key2.data();
    r.pair_counts = pair_counts.data();
  }
};

[[noreturn]] void Bad(const char *fn, const std::string &what) {
  throw duckdb::InvalidInputException(std::string(fn) + ": " + what);
}

void Unpack(const char *fn, const Value &v, bool nb, Unpacked &u) {
  const vector<Value> &f = duckdb::StructValue::GetChildren(v);
  if (f.size() != (nb ? 4u : 6u)) Bad(fn, "a ring value has " + std::string(nb ? "4" : "6") + " fields");
  const vector<Value> &lin = duckdb::ListValue::GetChildren(f[1]);
  const vector<Value> &quad = duckdb::ListValue::GetChildren(f[2]);
  const vector<Value> &lin_cat = duckdb::ListValue::GetChildren(f[3]);
  static const vector<Value> none;
  const vector<Value> &num_cat = nb ? none : duckdb::ListValue::GetChildren(f[4]);
  const vector<Value> &cat_cat = nb ? none : duckdb::ListValue::GetChildren(f[5]);
  cfb_result &r = u.r;
  r.kind = nb ? CFB_NB : CFB_TRIPLE;
  r.N = f[0].GetValue<int>();
  u.zero = lin.empty() && quad.empty() && lin_cat.empty() && num_cat.empty() && cat_cat.empty();
  const int n = (int)lin.size(), m = (int)lin_cat.size();
  r.n_num = n;
  r.n_cat = m;
  r.n_quad = nb ? n : (int64_t)n * (n + 1) / 2;
  if ((int64_t)quad.size() != r.n_quad) Bad(fn, "quad_num does not have the length lin_num implies");
  if (!nb && ((int64_t)num_cat.size() != (int64_t)n * m || (int64_t)cat_cat.size() != (int64_t)m * (m + 1) / 2))
    Bad(fn, "quad_num_cat / quad_cat do not have the lengths lin_num and lin_cat imply");
  for (auto &x : lin) u.lin.push_back(x.GetValue<float>());
  for (auto &x : quad) u.quad.push_back(x.GetValue<float>());
  // keys of column k: those of lin_cat[k] and of every quad_num_cat list of that column (the same set in a real triple)
  std::vector<std::map<int32_t, int64_t>> keys(m);
  for (int k = 0; k < m; k++) {
    for (auto &e : duckdb::ListValue::GetChildren(lin_cat[k])) {
      auto &kv = duckdb::StructValue::GetChildren(e);
      keys[k][kv[0].GetValue<int>()] = (int64_t)kv[1].GetValue<float>();
    }
    for (int i = 0; i < n && !nb; i++)
      for (auto &e : duckdb::ListValue::GetChildren(num_cat[(size_t)i * m + k]))
        keys[k].emplace(duckdb::StructValue::GetChildren(e)[0].GetValue<int>(), 0);
  }
  u.cat_offsets.assign(1, 0);
  std::vector<std::map<int32_t, int64_t>> pos(m);  // key -> index into cat_keys
  for (int k = 0; k < m; k++) {
    for (auto &kc : keys[k]) {
      pos[k][kc.first] = (int64_t)u.cat_keys.size();
      u.cat_keys.push_back(kc.first);
      u.cat_counts.push_back(kc.second);
    }
    u.cat_offsets.push_back((int64_t)u.cat_keys.size());
  }
  r.total_keys = (int64_t)u.cat_keys.size();
  u.pair_offsets.assign(1, 0);
  if (!nb) {
    u.numcat.assign((size_t)n * r.total_keys, 0.0);
    for (int i = 0; i < n; i++)
      for (int k = 0; k < m; k++)
        for (auto &e : duckdb::ListValue::GetChildren(num_cat[(size_t)i * m + k])) {
          auto &kv = duckdb::StructValue::GetChildren(e);
          u.numcat[(size_t)i * r.total_keys + pos[k][kv[0].GetValue<int>()]] = kv[1].GetValue<float>();
        }
    r.n_pair_lists = (int64_t)cat_cat.size();
    for (auto &l : cat_cat) {
      std::map<std::pair<int32_t, int32_t>, int64_t> pairs;
      for (auto &e : duckdb::ListValue::GetChildren(l)) {
        auto &kv = duckdb::StructValue::GetChildren(e);
        pairs[{kv[0].GetValue<int>(), kv[1].GetValue<int>()}] = (int64_t)kv[2].GetValue<float>();
      }
      for (auto &pc : pairs) {
        u.pair_key1.push_back(pc.first.first);
        u.pair_key2.push_back(pc.first.second);
        u.pair_counts.push_back(pc.second);
      }
      u.pair_offsets.push_back((int64_t)u.pair_key1.size());
    }
  }
  u.Bind();
}

// the ring's zero in the shape of `like` (no keys)
void ZeroLike(const Unpacked &like, Unpacked &z) {
  z.r = like.r;
  z.r.N = 0;
  z.r.total_keys = 0;
  z.lin.assign(like.lin.size(), 0.0);
  z.quad.assign(like.quad.size(), 0.0);
  z.cat_offsets.assign(like.r.n_cat + 1, 0);
  z.pair_offsets.assign(like.r.n_pair_lists + 1, 0);
  z.Bind();
}

Value KeyValueList(const cfb_result &r, int k, const double *sums, const int64_t *counts) {
  vector<Value> out;
  for (int64_t t = r.cat_offsets[k]; t < r.cat_offsets[k + 1]; t++) {
    duckdb::child_list_t<Value> kv;
    kv.emplace_back("key", Value((int32_t)r.cat_keys[t]));
    kv.emplace_back("value", Value(sums ? (float)sums[t] : (float)counts[t]));
    out.push_back(Value::STRUCT(std::move(kv)));
  }
  duckdb::child_list_t<duckdb::LogicalType> kv_t;
  kv_t.emplace_back("key", duckdb::LogicalType::INTEGER);
  kv_t.emplace_back("value", duckdb::LogicalType::FLOAT);
  return Value::LIST(duckdb::LogicalType::STRUCT(kv_t), std::move(out));
}

Value Pack(const cfb_result &r, bool nb) {
  duckdb::child_list_t<Value> s;
  s.emplace_back("N", Value((int32_t)r.N));
  vector<Value> lin, quad, lin_cat, num_cat, cat_cat;
  for (int i = 0; i < r.n_num; i++) lin.push_back(Value((float)r.lin[i]));
  for (int64_t i = 0; i < r.n_quad; i++) quad.push_back(Value((float)r.quad[i]));
  s.emplace_back("lin_num", Value::LIST(duckdb::LogicalType::FLOAT, std::move(lin)));
  s.emplace_back("quad_num", Value::LIST(duckdb::LogicalType::FLOAT, std::move(quad)));
  duckdb::child_list_t<duckdb::LogicalType> kv_t, kkv_t;
  kv_t.emplace_back("key", duckdb::LogicalType::INTEGER);
  kv_t.emplace_back("value", duckdb::LogicalType::FLOAT);
  const auto list_kv = duckdb::LogicalType::LIST(duckdb::LogicalType::STRUCT(kv_t));
  for (int k = 0; k < r.n_cat; k++) lin_cat.push_back(KeyValueList(r, k, nullptr, r.cat_counts));
  s.emplace_back("lin_cat", Value::LIST(list_kv, std::move(lin_cat)));
  if (nb) return Value::STRUCT(std::move(s));
  for (int i = 0; i < r.n_num; i++)
    for (int k = 0; k < r.n_cat; k++) num_cat.push_back(KeyValueList(r, k, r.numcat_sums + (size_t)i * r.total_keys, nullptr));
  s.emplace_back("quad_num_cat", Value::LIST(list_kv, std::move(num_cat)));
  kkv_t.emplace_back("key1", duckdb::LogicalType::INTEGER);
  kkv_t.emplace_back("key2", duckdb::LogicalType::INTEGER);
  kkv_t.emplace_back("value", duckdb::LogicalType::FLOAT);
  for (int64_t p = 0; p < r.n_pair_lists; p++) {
    vector<Value> l;
    for (int64_t t = r.pair_offsets[p]; t < r.pair_offsets[p + 1]; t++) {
      duckdb::child_list_t<Value> e;
      e.emplace_back("key1", Value((int32_t)r.pair_key1[t]));
      e.emplace_back("key2", Value((int32_t)r.pair_key2[t]));
      e.emplace_back("value", Value((float)r.pair_counts[t]));
      l.push_back(Value::STRUCT(std::move(e)));
    }
    cat_cat.push_back(Value::LIST(duckdb::LogicalType::STRUCT(kkv_t), std::move(l)));
  }
  s.emplace_back("quad_cat", Value::LIST(duckdb::LogicalType::LIST(duckdb::LogicalType::STRUCT(kkv_t)), std::move(cat_cat)));
  return Value::STRUCT(std::move(s));
}

Value Ring(const char *fn, const Value &a, const Value &b, bool nb, int sign) {
  Unpacked ua, ub, z;
  Unpack(fn, a, nb, ua);
  Unpack(fn, b, nb, ub);
  const cfb_result *pa = &ua.r, *pb = &ub.r;
  if (ua.zero && !ub.zero) {
    ZeroLike(ub, z);
    z.r.N = ua.r.N;
    pa = &z.r;
  } else if (ub.zero && !ua.zero) {
    ZeroLike(ua, z);
    z.r.N = ub.r.N;
    pb = &z.r;
  }
  cfb_result out;
  if (cfb_result_combine(pa, pb, sign, CFB_COMBINE_KEEP_ZERO_KEYS, &out) != CFB_OK) Bad(fn, cfb_last_error());
  Value v = Pack(out, nb);
  cfb_result_free(&out);
  return v;
}

}  // namespace

namespace Triple {
// replaces imputation/triple/sum.cpp:69-209
Value sum_triple(const Value &triple_1, const Value &triple_2) { return Ring("sum_triple", triple_1, triple_2, false, +1); }
// replaces imputation/triple/sub.cpp:71-216
Value subtract_triple(Value &triple_1, Value &triple_2) { return Ring("subtract_triple", triple_1, triple_2, false, -1); }
// replaces imputation/triple/sum_nb.cpp:38-82
Value sum_nb_triple(const Value &triple_1, const Value &triple_2) { return Ring("sum_nb_triple", triple_1, triple_2, true, +1); }
}  // namespace Triple
