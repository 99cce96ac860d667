// cofactor_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see cofactor_oracle.h).
//
// CPU restatement of the ring ("triple") sum aggregates of eddbase/duckdb-imputation.
// Citations are file:line under /root/reference/duckdb_extension/src.
//
//   ORC_EXACT    int64 counts + fp64 sums in one pass (the parity target at scale).
//   ORC_FAITHFUL the reference's arithmetic: fp32 sums, fp32 counts, the update's own
//                loop nest (all of lin, then pair-by-pair column sweeps, then per-row
//                map lookups), 2048-row chunks, a per-row state-pointer vector, T worker
//                threads with private states, then combine, then finalize.
#include "cofactor_oracle.h"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <map>
#include <thread>
#include <utility>
#include <vector>

namespace {

constexpr size_t kChunk = 2048;  // STANDARD_VECTOR_SIZE: rows per update call

thread_local double g_last_seconds = 0.0;

template <class T>
T *dup(const std::vector<T> &v) {
  T *p = static_cast<T *>(std::malloc(std::max<size_t>(1, v.size()) * sizeof(T)));
  if (!v.empty()) std::memcpy(p, v.data(), v.size() * sizeof(T));
  return p;
}

// ------------------------------------------------------------------ exact state
struct ExactState {
  int64_t N = 0;
  std::vector<double> lin, quad;
  std::vector<std::map<int32_t, std::vector<double>>> numcat;  // key -> [count, sum x_1..x_n]
  std::vector<std::map<std::pair<int32_t, int32_t>, int64_t>> catcat;
};

// ------------------------------------------------------------- faithful state
// Mirrors Triple::SumState (sum_state.h:14-28): int count, one float block holding
// lin then quad, an array of m ordered maps key -> [count, sums...] and an array of
// m(m+1)/2 ordered maps (key1,key2) -> float count.
struct RefState {
  int count = 0;
  int n = 0, m = 0;
  bool nb = false;
  bool shaped = false;
  std::vector<float> lin, quad;
  std::vector<std::map<int, std::vector<float>>> numcat;
  std::vector<std::map<std::pair<int, int>, float>> catcat;
};

struct Input {
  int kind, n, m;
  const float *const *num;
  const int32_t *const *cat;
  const int32_t *group;
  const uint32_t *sel;
};

inline size_t src_row(const Input &in, size_t r) { return in.sel ? in.sel[r] : r; }

void shape_ref(RefState &s, const Input &in) {
  // lazy allocation on the first row a state sees (sum_no_lift.cpp:96-116,
  // sum_to_nb_agg.cpp:75-95)
  s.n = in.n;
  s.m = in.m;
  s.nb = in.kind == ORC_NB;
  s.lin.assign(in.n, 0.f);
  s.quad.assign(s.nb ? in.n : in.n * (in.n + 1) / 2, 0.f);
  s.numcat.assign(in.m, {});
  s.catcat.assign(s.nb ? 0 : in.m * (in.m + 1) / 2, {});
  s.shaped = true;
}

// One update call over rows [lo, hi) (hi-lo <= 2048); `st` = per-row state pointers.
void ref_update_chunk(const Input &in, size_t lo, size_t hi, RefState **st) {
  const size_t cnt = hi - lo;
  const int n = in.n, m = in.m;
  // count (sum_no_lift.cpp:83-86)
  for (size_t r = 0; r < cnt; r++) st[r]->count += 1;
  // shape + lin (sum_no_lift.cpp:91-122)
  for (size_t r = 0; r < cnt; r++) {
    RefState &s = *st[r];
    if (!s.shaped) shape_ref(s, in);
    const size_t row = src_row(in, lo + r);
    for (int k = 0; k < n; k++) s.lin[k] += in.num[k][row];
  }
  if (in.kind == ORC_TRIPLE) {
    // numeric pairs j<=k, one column sweep per pair (sum_no_lift.cpp:128-147)
    int p = 0;
    for (int j = 0; j < n; j++)
      for (int k = j; k < n; k++, p++)
        for (size_t r = 0; r < cnt; r++) {
          const size_t row = src_row(in, lo + r);
          st[r]->quad[p] += in.num[j][row] * in.num[k][row];
        }
    // per row, per categorical column: payload [count, x_1..x_n] (sum_no_lift.cpp:158-190)
    for (size_t r = 0; r < cnt; r++) {
      RefState &s = *st[r];
      const size_t row = src_row(in, lo + r);
      for (int c = 0; c < m; c++) {
        auto &mp = s.numcat[c];
        const int key = in.cat[c][row];
        auto it = mp.find(key);
        if (it == mp.end()) {
          std::vector<float> pay(n + 1);
          pay[0] = 1.f;
          for (int k = 0; k < n; k++) pay[k + 1] = in.num[k][row];
          mp.emplace(key, std::move(pay));
        } else {
          for (int k = 0; k < n; k++) it->second[k + 1] += in.num[k][row];
          it->second[0] += 1.f;
        }
      }
    }
    // categorical pairs j<=k, diagonal included (sum_no_lift.cpp:195-214)
    p = 0;
    for (int j = 0; j < m; j++)
      for (int k = j; k < m; k++, p++)
        for (size_t r = 0; r < cnt; r++) {
          const size_t row = src_row(in, lo + r);
          auto &mp = st[r]->catcat[p];
          const std::pair<int, int> key(in.cat[j][row], in.cat[k][row]);
          auto it = mp.find(key);
          if (it == mp.end())
            mp.emplace(key, 1.f);
          else
            it->second += 1.f;
        }
  } else {
    // NB: diagonal only (sum_to_nb_agg.cpp:104-117) and key counts only (:124-145)
    for (int j = 0; j < n; j++)
      for (size_t r = 0; r < cnt; r++) {
        const float v = in.num[j][src_row(in, lo + r)];
        st[r]->quad[j] += v * v;
      }
    for (size_t r = 0; r < cnt; r++) {
      RefState &s = *st[r];
      const size_t row = src_row(in, lo + r);
      for (int c = 0; c < m; c++) {
        auto &mp = s.numcat[c];
        const int key = in.cat[c][row];
        auto it = mp.find(key);
        if (it == mp.end())
          mp.emplace(key, std::vector<float>(1, 1.f));
        else
          it->second[0] += 1.f;
      }
    }
  }
}

// dst += src (sum_state.cpp:23-112)
void ref_combine(RefState &dst, const RefState &src) {
  dst.count += src.count;
  if (!src.shaped) return;  // the reference would dereference nullptr here; never happens
  if (!dst.shaped) {
    dst.n = src.n;
    dst.m = src.m;
    dst.nb = src.nb;
    dst.lin.assign(src.lin.size(), 0.f);
    dst.quad.assign(src.quad.size(), 0.f);
    dst.numcat.assign(src.numcat.size(), {});
    dst.catcat.assign(src.catcat.size(), {});
    dst.shaped = true;
  }
  for (size_t k = 0; k < dst.lin.size(); k++) dst.lin[k] += src.lin[k];
  for (size_t k = 0; k < dst.quad.size(); k++) dst.quad[k] += src.quad[k];
  for (size_t c = 0; c < dst.numcat.size(); c++)
    for (const auto &kv : src.numcat[c]) {
      auto it = dst.numcat[c].find(kv.first);
      if (it == dst.numcat[c].end())
        dst.numcat[c].emplace(kv.first, kv.second);
      else
        for (size_t l = 0; l < kv.second.size(); l++) it->second[l] += kv.second[l];
    }
  for (size_t p = 0; p < dst.catcat.size(); p++)
    for (const auto &kv : src.catcat[p]) {
      auto it = dst.catcat[p].find(kv.first);
      if (it == dst.catcat[p].end())
        dst.catcat[p].emplace(kv.first, kv.second);
      else
        it->second += kv.second;
    }
}

// Canonical flat result in SumStateFinalize's order (sum_state.cpp:132-461): keys ascending
// per categorical column (std::map iteration), quad_num_cat for numeric i over the same key
// sequence, pair lists in (k<=l) order with entries ascending by (key1,key2).
template <class NumCat, class CatCat, class F>
void emit(int kind, int n, int m, int64_t N, const F *lin, size_t n_lin, const F *quad, size_t n_quad,
          const NumCat &numcat, const CatCat &catcat, orc_result *out) {
  std::memset(out, 0, sizeof(*out));
  out->kind = kind;
  out->n_num = n;
  out->n_cat = m;
  out->N = N;
  out->n_quad = (kind == ORC_NB) ? n : (int64_t)n * (n + 1) / 2;
  std::vector<double> l(n, 0.0), q(out->n_quad, 0.0);
  for (size_t i = 0; i < n_lin; i++) l[i] = (double)lin[i];
  for (size_t i = 0; i < n_quad; i++) q[i] = (double)quad[i];
  out->lin = dup(l);
  out->quad = dup(q);
  std::vector<int64_t> offs(m + 1, 0), counts;
  std::vector<int32_t> keys;
  for (int c = 0; c < m; c++) {
    if ((size_t)c < numcat.size())
      for (const auto &kv : numcat[c]) {
        keys.push_back(kv.first);
        counts.push_back((int64_t)kv.second[0]);
      }
    offs[c + 1] = (int64_t)keys.size();
  }
  out->total_keys = (int64_t)keys.size();
  out->cat_offsets = dup(offs);
  out->cat_keys = dup(keys);
  out->cat_counts = dup(counts);
  if (kind == ORC_TRIPLE) {
    std::vector<double> nc((size_t)n * keys.size(), 0.0);
    size_t t = 0;
    for (int c = 0; c < m && (size_t)c < numcat.size(); c++)
      for (const auto &kv : numcat[c]) {
        for (int i = 0; i < n; i++) nc[(size_t)i * keys.size() + t] = (double)kv.second[i + 1];
        t++;
      }
    out->numcat_sums = dup(nc);
    out->n_pair_lists = (int64_t)m * (m + 1) / 2;
    std::vector<int64_t> po(out->n_pair_lists + 1, 0), pc;
    std::vector<int32_t> k1, k2;
    for (int64_t p = 0; p < out->n_pair_lists; p++) {
      if ((size_t)p < catcat.size())
        for (const auto &kv : catcat[p]) {
          k1.push_back(kv.first.first);
          k2.push_back(kv.first.second);
          pc.push_back((int64_t)kv.second);
        }
      po[p + 1] = (int64_t)k1.size();
    }
    out->pair_offsets = dup(po);
    out->pair_key1 = dup(k1);
    out->pair_key2 = dup(k2);
    out->pair_counts = dup(pc);
  } else {
    out->n_pair_lists = 0;
    out->pair_offsets = dup(std::vector<int64_t>(1, 0));
  }
}

void emit_ref(const RefState &s, int kind, int n, int m, orc_result *out) {
  emit(kind, n, m, (int64_t)s.count, s.lin.data(), s.lin.size(), s.quad.data(), s.quad.size(), s.numcat,
       s.catcat, out);
}

int run_exact(const Input &in, int n_groups, size_t rows, orc_result *out) {
  const int n = in.n, m = in.m;
  const bool nb = in.kind == ORC_NB;
  std::vector<ExactState> st(n_groups);
  for (auto &s : st) {
    s.lin.assign(n, 0.0);
    s.quad.assign(nb ? n : n * (n + 1) / 2, 0.0);
    s.numcat.resize(m);
    s.catcat.resize(nb ? 0 : m * (m + 1) / 2);
  }
  std::vector<double> x(n);
  for (size_t r = 0; r < rows; r++) {
    const size_t row = src_row(in, r);
    const int g = in.group ? in.group[row] : 0;
    if (g < 0 || g >= n_groups) return -1;
    ExactState &s = st[g];
    s.N++;
    for (int k = 0; k < n; k++) {
      x[k] = (double)in.num[k][row];
      s.lin[k] += x[k];
    }
    if (nb) {
      for (int k = 0; k < n; k++) s.quad[k] += x[k] * x[k];
    } else {
      int p = 0;
      for (int j = 0; j < n; j++)
        for (int k = j; k < n; k++, p++) s.quad[p] += x[j] * x[k];
    }
    for (int c = 0; c < m; c++) {
      auto &pay = s.numcat[c][in.cat[c][row]];
      if (pay.empty()) pay.assign(nb ? 1 : n + 1, 0.0);
      pay[0] += 1.0;
      if (!nb)
        for (int k = 0; k < n; k++) pay[k + 1] += x[k];
    }
    if (!nb) {
      int p = 0;
      for (int j = 0; j < m; j++)
        for (int k = j; k < m; k++, p++) s.catcat[p][{in.cat[j][row], in.cat[k][row]}] += 1;
    }
  }
  for (int g = 0; g < n_groups; g++)
    emit(in.kind, n, m, st[g].N, st[g].lin.data(), st[g].lin.size(), st[g].quad.data(),
         st[g].quad.size(), st[g].numcat, st[g].catcat, &out[g]);
  return 0;
}

// With a filter the scan still walks the table in physical 2048-row chunks and every chunk
// reaches update() with a selection on top (count = rows that pass): chunks and thread morsels
// are cut in PHYSICAL rows, exactly as the replay host does.
int run_faithful_filtered(const Input &in, int n_groups, size_t n_sel, size_t table_rows, int threads, orc_result *out) {
  if (threads < 1) threads = 1;
  if (n_sel && in.sel[n_sel - 1] >= table_rows) table_rows = (size_t)in.sel[n_sel - 1] + 1;
  const size_t n_chunks = (table_rows + kChunk - 1) / kChunk;
  if ((size_t)threads > std::max<size_t>(1, n_chunks)) threads = (int)std::max<size_t>(1, n_chunks);
  std::vector<std::vector<RefState>> local(threads, std::vector<RefState>(n_groups));
  int bad = 0;
  auto worker = [&](int t) {
    const size_t c_lo = n_chunks * t / threads, c_hi = n_chunks * (t + 1) / threads;
    std::vector<RefState *> st(kChunk);
    size_t pos = std::lower_bound(in.sel, in.sel + n_sel, (uint32_t)(c_lo * kChunk)) - in.sel;
    for (size_t c = c_lo; c < c_hi; c++) {
      const size_t hi = std::min(table_rows, (c + 1) * kChunk);
      const size_t first = pos;
      while (pos < n_sel && in.sel[pos] < hi) {
        const int g = in.group ? in.group[in.sel[pos]] : 0;
        if (g < 0 || g >= n_groups) {
          bad = 1;
          return;
        }
        st[pos - first] = &local[t][g];
        pos++;
      }
      if (pos > first) ref_update_chunk(in, first, pos, st.data());  // rows index the selection
    }
  };
  if (threads == 1) {
    worker(0);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) pool.emplace_back(worker, t);
    for (auto &th : pool) th.join();
  }
  if (bad) return -1;
  for (int g = 0; g < n_groups; g++) {
    RefState total;
    for (int t = 0; t < threads; t++) ref_combine(total, local[t][g]);
    emit_ref(total, in.kind, in.n, in.m, &out[g]);
  }
  return 0;
}

int run_faithful(const Input &in, int n_groups, size_t rows, int threads, orc_result *out) {
  if (threads < 1) threads = 1;
  const size_t n_chunks = (rows + kChunk - 1) / kChunk;
  if ((size_t)threads > std::max<size_t>(1, n_chunks)) threads = (int)std::max<size_t>(1, n_chunks);
  // one private state per (thread, group): DuckDB's thread-local hash tables
  std::vector<std::vector<RefState>> local(threads, std::vector<RefState>(n_groups));
  int bad = 0;
  auto worker = [&](int t) {
    const size_t c_lo = n_chunks * t / threads, c_hi = n_chunks * (t + 1) / threads;
    std::vector<RefState *> st(kChunk);
    for (size_t c = c_lo; c < c_hi; c++) {
      const size_t lo = c * kChunk, hi = std::min(rows, lo + kChunk);
      for (size_t r = lo; r < hi; r++) {
        const int g = in.group ? in.group[src_row(in, r)] : 0;
        if (g < 0 || g >= n_groups) {
          bad = 1;
          return;
        }
        st[r - lo] = &local[t][g];
      }
      ref_update_chunk(in, lo, hi, st.data());
    }
  };
  if (threads == 1) {
    worker(0);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) pool.emplace_back(worker, t);
    for (auto &th : pool) th.join();
  }
  if (bad) return -1;
  for (int g = 0; g < n_groups; g++) {
    RefState total;
    for (int t = 0; t < threads; t++) ref_combine(total, local[t][g]);
    emit_ref(total, in.kind, in.n, in.m, &out[g]);
  }
  return 0;
}

}  // namespace

extern "C" {

int orc_aggregate(int kind, int mode, int n_num, int n_cat, const float *const *num_cols,
                  const int32_t *const *cat_cols, const int32_t *group, int n_groups,
                  const uint32_t *sel, size_t rows, size_t table_rows, int threads, orc_result *out) {
  if ((kind != ORC_TRIPLE && kind != ORC_NB) || n_num < 0 || n_cat < 0 || n_groups < 1 || !out) return -1;
  Input in{kind, n_num, n_cat, num_cols, cat_cols, group, sel};
  const auto t0 = std::chrono::steady_clock::now();
  const int rc = (mode == ORC_EXACT) ? run_exact(in, n_groups, rows, out)
                 : sel              ? run_faithful_filtered(in, n_groups, rows, table_rows, threads, out)
                                    : run_faithful(in, n_groups, rows, threads, out);
  g_last_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return rc;
}

double orc_last_seconds(void) { return g_last_seconds; }

// to_cofactor / to_nb_agg followed by sum_triple / sum_nb_agg.  A lifted row is the
// triple N=1, lin=[x], quad=[x_j*x_k] (lift.cpp:119-136; NB: x_j^2, lift_to_nb_agg.cpp:108-117),
// lin_cat[c]={key:1}, quad_num_cat[j*m+c]={key:x_j} (lift.cpp:157-176),
// quad_cat[(j<=k)]={(c_j,c_k):1} (lift.cpp:199-219).  Sum adds them field by field in fp32
// (sum.cpp:86-260).  The products are formed in fp32 by the lift, then added.
int orc_sum_of_lifted(int kind, int n_num, int n_cat, const float *const *num_cols,
                      const int32_t *const *cat_cols, const int32_t *group, int n_groups,
                      size_t rows, orc_result *out) {
  if ((kind != ORC_TRIPLE && kind != ORC_NB) || n_groups < 1 || !out) return -1;
  Input in{kind, n_num, n_cat, num_cols, cat_cols, group, nullptr};
  std::vector<RefState> st(n_groups);
  const int n = n_num, m = n_cat;
  for (size_t row = 0; row < rows; row++) {
    const int g = group ? group[row] : 0;
    if (g < 0 || g >= n_groups) return -1;
    RefState &s = st[g];
    s.count += 1;  // N of a lifted row is 1
    if (!s.shaped) shape_ref(s, in);
    for (int k = 0; k < n; k++) s.lin[k] += num_cols[k][row];
    if (kind == ORC_TRIPLE) {
      int p = 0;
      for (int j = 0; j < n; j++)
        for (int k = j; k < n; k++, p++) {
          const float prod = num_cols[j][row] * num_cols[k][row];
          s.quad[p] += prod;
        }
    } else {
      for (int j = 0; j < n; j++) {
        const float prod = num_cols[j][row] * num_cols[j][row];
        s.quad[j] += prod;
      }
    }
    for (int c = 0; c < m; c++) {
      auto &mp = s.numcat[c];
      const int key = cat_cols[c][row];
      auto it = mp.find(key);
      if (it == mp.end()) {
        std::vector<float> pay(kind == ORC_NB ? 1 : n + 1);
        pay[0] = 1.f;
        if (kind == ORC_TRIPLE)
          for (int k = 0; k < n; k++) pay[k + 1] = num_cols[k][row];
        mp.emplace(key, std::move(pay));
      } else {
        it->second[0] += 1.f;
        if (kind == ORC_TRIPLE)
          for (int k = 0; k < n; k++) it->second[k + 1] += num_cols[k][row];
      }
    }
    if (kind == ORC_TRIPLE) {
      int p = 0;
      for (int j = 0; j < m; j++)
        for (int k = j; k < m; k++, p++) {
          const std::pair<int, int> key(cat_cols[j][row], cat_cols[k][row]);
          auto it = s.catcat[p].find(key);
          if (it == s.catcat[p].end())
            s.catcat[p].emplace(key, 1.f);
          else
            it->second += 1.f;
        }
    }
  }
  for (int g = 0; g < n_groups; g++) emit_ref(st[g], kind, n, m, &out[g]);
  return 0;
}

int orc_result_add(const orc_result *a, const orc_result *b, orc_result *out) {
  if (!a || !b || !out || a->kind != b->kind || a->n_num != b->n_num || a->n_cat != b->n_cat) return -1;
  const int n = a->n_num, m = a->n_cat;
  ExactState s;
  s.N = a->N + b->N;
  s.lin.assign(n, 0.0);
  s.quad.assign(a->n_quad, 0.0);
  s.numcat.resize(m);
  s.catcat.resize(a->n_pair_lists);
  for (const orc_result *r : {a, b}) {
    for (int i = 0; i < n; i++) s.lin[i] += r->lin[i];
    for (int64_t i = 0; i < r->n_quad; i++) s.quad[i] += r->quad[i];
    for (int c = 0; c < m; c++)
      for (int64_t t = r->cat_offsets[c]; t < r->cat_offsets[c + 1]; t++) {
        auto &pay = s.numcat[c][r->cat_keys[t]];
        if (pay.empty()) pay.assign(a->kind == ORC_NB ? 1 : n + 1, 0.0);
        pay[0] += (double)r->cat_counts[t];
        if (a->kind == ORC_TRIPLE)
          for (int i = 0; i < n; i++) pay[i + 1] += r->numcat_sums[(size_t)i * r->total_keys + t];
      }
    for (int64_t p = 0; p < r->n_pair_lists; p++)
      for (int64_t t = r->pair_offsets[p]; t < r->pair_offsets[p + 1]; t++)
        s.catcat[p][{r->pair_key1[t], r->pair_key2[t]}] += r->pair_counts[t];
  }
  emit(a->kind, n, m, s.N, s.lin.data(), s.lin.size(), s.quad.data(), s.quad.size(), s.numcat, s.catcat, out);
  return 0;
}

void orc_result_free(orc_result *r) {
  if (!r) return;
  std::free(r->lin);
  std::free(r->quad);
  std::free(r->cat_offsets);
  std::free(r->cat_keys);
  std::free(r->cat_counts);
  std::free(r->numcat_sums);
  std::free(r->pair_offsets);
  std::free(r->pair_key1);
  std::free(r->pair_key2);
  std::free(r->pair_counts);
  std::memset(r, 0, sizeof(*r));
}

}  // extern "C"
