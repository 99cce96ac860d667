// triple_glue.cpp -- DuckDB aggregate callbacks -> C ABI (include/cofactor_b200.h).
//
//   update    Triple::SumNoLift / Triple::sum_to_nb_agg   (reference: sum_no_lift.cpp:53-216,
//             sum_to_nb_agg.cpp:39-146): read the chunk through UnifiedVectorFormat, route
//             rows to their states, hand column base pointers + selection vectors to
//             cfb_ctx_append (pinned staging -> cudaMemcpyAsync -> kernels).
//   combine   Triple::SumStateCombine (sum_state.cpp:10-114): adopt or cfb_ctx_combine.
//   finalize  Triple::SumStateFinalize (sum_state.cpp:116-464): cfb_ctx_finalize, then write
//             the nested STRUCT vector in the reference's layout.
// Errors of the C ABI become duckdb::InternalException / InvalidInputException (never abort).
#include "triple_glue.h"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../../include/cofactor_b200.h"

namespace Triple {

namespace {

void Check(int rc) {
  if (rc == CFB_OK) return;
  const std::string msg = cfb_last_error();
  if (rc == CFB_ERR_INVALID || rc == CFB_ERR_DOMAIN) throw duckdb::InvalidInputException(msg);
  throw duckdb::InternalException(msg);
}

// Device of a new state: CFB_DEVICE=<k> pins every state to device k (one process per GPU);
// CFB_DEVICES=all spreads states over the visible devices (one DuckDB worker thread keeps
// feeding the same state, so a state's device never changes); default: device 0.
int PickDevice() {
  if (const char *one = getenv("CFB_DEVICE")) return atoi(one);
  static const int n_dev = [] {
    const char *e = getenv("CFB_DEVICES");
    if (!e || strcmp(e, "all") != 0) return 1;
    return cfb_device_count() > 0 ? cfb_device_count() : 1;
  }();
  static std::atomic<unsigned> next{0};
  return n_dev == 1 ? 0 : (int)(next.fetch_add(1) % (unsigned)n_dev);
}

Arena *NewArena(int kind, int n, int m, int capacity) {
  Arena *a = new Arena();
  a->kind = kind;
  a->n = n;
  a->m = m;
  a->capacity = capacity;
  const int rc = cfb_ctx_create(PickDevice(), kind, n, m, capacity, &a->ctx);
  if (rc != CFB_OK) {
    delete a;
    Check(rc);
  }
  return a;
}

// Slots of a thread's FIRST shared arena: what the GROUP BY kernel keeps in shared memory
// (group_kernel.cuh); with pair tables the per-slot state is large, so few slots.  When an arena is
// full the next one is twice as large (up to kMaxCapacity), so G groups need O(log G) arenas.
int FirstCapacity(int kind, int n, int m) {
  if (kind == CFB_TRIPLE && m >= 2) return 8;
  const int entries = 1 + n + (kind == CFB_TRIPLE ? n * (n + 1) / 2 : n);
  return entries <= 128 ? 32 : 16;
}
int NextCapacity(int kind, int m, int prev) {
  const int max_cap = (kind == CFB_TRIPLE && m >= 2) ? 8 : 4096;
  return prev * 2 < max_cap ? prev * 2 : max_cap;
}

// The arenas this thread is currently filling, one per aggregate shape.
struct OpenArenas {
  std::vector<Arena *> open;
  ~OpenArenas() {
    for (Arena *a : open) a->Release();
  }
  Arena *Get(int kind, int n, int m) {
    for (size_t i = 0; i < open.size(); i++) {
      Arena *a = open[i];
      if (a->kind == kind && a->n == n && a->m == m) {
        if (a->next_slot < a->capacity) return a;
        const int cap = NextCapacity(kind, m, a->capacity);
        a->Release();  // full: the states keep it alive; start the next, larger one
        open[i] = NewArena(kind, n, m, cap);
        open[i]->refs.fetch_add(1);
        return open[i];
      }
    }
    Arena *a = NewArena(kind, n, m, FirstCapacity(kind, n, m));
    a->refs.fetch_add(1);
    open.push_back(a);
    return a;
  }
};
thread_local OpenArenas t_open;

duckdb::LogicalType KeyValueList() {
  duckdb::child_list_t<duckdb::LogicalType> kv;
  kv.emplace_back("key", duckdb::LogicalType::INTEGER);
  kv.emplace_back("value", duckdb::LogicalType::FLOAT);
  return duckdb::LogicalType::LIST(duckdb::LogicalType::LIST(duckdb::LogicalType::STRUCT(kv)));
}

// STRUCT(N, lin_agg, quad_agg, lin_cat[, quad_num_cat, quad_cat])  -- sum_no_lift.cpp:21-44,
// sum_to_nb_agg.cpp:18-30
duckdb::LogicalType ResultType(bool nb) {
  duckdb::child_list_t<duckdb::LogicalType> f;
  f.emplace_back("N", duckdb::LogicalType::INTEGER);
  f.emplace_back("lin_agg", duckdb::LogicalType::LIST(duckdb::LogicalType::FLOAT));
  f.emplace_back("quad_agg", duckdb::LogicalType::LIST(duckdb::LogicalType::FLOAT));
  f.emplace_back("lin_cat", KeyValueList());
  if (!nb) {
    f.emplace_back("quad_num_cat", KeyValueList());
    duckdb::child_list_t<duckdb::LogicalType> kkv;
    kkv.emplace_back("key1", duckdb::LogicalType::INTEGER);
    kkv.emplace_back("key2", duckdb::LogicalType::INTEGER);
    kkv.emplace_back("value", duckdb::LogicalType::FLOAT);
    f.emplace_back("quad_cat", duckdb::LogicalType::LIST(duckdb::LogicalType::LIST(duckdb::LogicalType::STRUCT(kkv))));
  }
  return duckdb::LogicalType::STRUCT(f);
}

// The chunk's columns as base pointers + selection vectors (UnifiedVectorFormat, sum_no_lift.cpp:66-73):
// FLOAT (and DOUBLE, like the reference) columns are numeric, everything else categorical; numeric
// columns come first (README.md:126).
struct ChunkColumns {
  const float *num[CFB_MAX_NUM];
  const uint32_t *num_sel[CFB_MAX_NUM];
  const int32_t *cat[CFB_MAX_CAT];
  const uint32_t *cat_sel[CFB_MAX_CAT];
  duckdb::UnifiedVectorFormat fmt[CFB_MAX_NUM + CFB_MAX_CAT];
  int n = 0, m = 0;
  ChunkColumns(duckdb::Vector inputs[], idx_t cols, idx_t count) {
    if (cols > CFB_MAX_NUM + CFB_MAX_CAT) throw duckdb::InvalidInputException("too many columns for a ring aggregate");
    for (idx_t j = 0; j < cols; j++) {
      inputs[j].ToUnifiedFormat(count, fmt[j]);
      const auto &t = inputs[j].GetType();
      if (t == duckdb::LogicalType::FLOAT || t == duckdb::LogicalType::DOUBLE) {
        if (m) throw duckdb::InvalidInputException("numeric columns must precede categorical columns");
        if (n == CFB_MAX_NUM) throw duckdb::InvalidInputException("too many numeric columns");
        num[n] = duckdb::UnifiedVectorFormat::GetData<float>(fmt[j]);
        num_sel[n++] = fmt[j].sel->data();
      } else {
        if (m == CFB_MAX_CAT) throw duckdb::InvalidInputException("too many categorical columns");
        cat[m] = duckdb::UnifiedVectorFormat::GetData<int32_t>(fmt[j]);
        cat_sel[m++] = fmt[j].sel->data();
      }
    }
  }
};

void AppendPrivate(SumState &state, int kind, const ChunkColumns &c, idx_t count) {
  // a state that is fed alone gets a private one-slot context: the ungrouped kernels apply
  cfb_ctx *ctx = PrivateContext(state, kind, c.n, c.m);
  std::lock_guard<std::mutex> g(state.arena->mu);
  Check(cfb_ctx_append(ctx, c.num, c.num_sel, c.cat, c.cat_sel, nullptr, count));
}

void Update(int kind, duckdb::Vector inputs[], idx_t cols, duckdb::Vector &state_vector, idx_t count) {
  if (count == 0) return;
  duckdb::UnifiedVectorFormat sdata;
  state_vector.ToUnifiedFormat(count, sdata);
  auto states = (SumState **)sdata.data;
  const ChunkColumns c(inputs, cols, count);
  const int n = c.n, m = c.m;
  // Ungrouped aggregates (and single-group chunks) send every row to one state.
  SumState *first = states[sdata.sel->get_index(0)];
  bool uniform = true;
  if (state_vector.GetVectorType() != duckdb::VectorType::CONSTANT_VECTOR)
    for (idx_t r = 1; r < count && uniform; r++) uniform = states[sdata.sel->get_index(r)] == first;
  if (uniform && (!first->arena || first->arena->capacity == 1)) {
    AppendPrivate(*first, kind, c, count);
    return;
  }
  // GROUP BY: new states get a slot in this thread's open arena; the chunk is shipped once per arena
  // present in it (normally one) with a slot id per row -- rows of other arenas are marked -1 and
  // dropped on the device (the states[sdata.sel->get_index(j)] indirection of the reference).
  static thread_local std::vector<uint32_t> slots;
  slots.resize(count);
  static thread_local std::vector<Arena *> seen;
  seen.clear();
  for (idx_t r = 0; r < count; r++) {
    SumState *s = states[sdata.sel->get_index(r)];
    if (!s->arena) AssignSlot(*s, kind, n, m);
    size_t b = 0;
    while (b < seen.size() && seen[b] != s->arena) b++;
    if (b == seen.size()) seen.push_back(s->arena);
  }
  for (size_t b = 0; b < seen.size(); b++) {
    Arena *a = seen[b];
    if (a->kind != kind || a->n != n || a->m != m) throw duckdb::InvalidInputException("ring aggregate: state shape changed");
    for (idx_t r = 0; r < count; r++) {
      const SumState *s = states[sdata.sel->get_index(r)];
      slots[r] = s->arena == a ? (uint32_t)s->slot : 0xFFFFFFFFu;  // -1 as int32: row not for this arena
    }
    std::lock_guard<std::mutex> g(a->mu);
    Check(cfb_ctx_append(a->ctx, c.num, c.num_sel, c.cat, c.cat_sel, slots.data(), count));
  }
}

void SimpleUpdate(int kind, duckdb::Vector inputs[], idx_t cols, duckdb::data_ptr_t state, idx_t count) {
  if (count == 0) return;
  SumState &s = *reinterpret_cast<SumState *>(state);
  const ChunkColumns c(inputs, cols, count);
  if (!s.arena || s.arena->capacity == 1) {
    AppendPrivate(s, kind, c, count);
    return;
  }
  // (a state that already lives in a GROUP BY arena: every row goes to its slot)
  static thread_local std::vector<uint32_t> slots;
  slots.assign(count, (uint32_t)s.slot);
  std::lock_guard<std::mutex> g(s.arena->mu);
  Check(cfb_ctx_append(s.arena->ctx, c.num, c.num_sel, c.cat, c.cat_sel, slots.data(), count));
}

}  // namespace

cfb_ctx *PrivateContext(SumState &state, int kind, int n_num, int n_cat) {
  if (!state.arena) {
    state.arena = NewArena(kind, n_num, n_cat, 1);  // lazy shape, sum_no_lift.cpp:96-116
    state.arena->refs.fetch_add(1);
    state.slot = 0;
  }
  return state.arena->ctx;
}

void AssignSlot(SumState &state, int kind, int n_num, int n_cat) {
  if (state.arena) return;
  Arena *a = t_open.Get(kind, n_num, n_cat);
  state.arena = a;
  state.slot = a->next_slot++;
  a->refs.fetch_add(1);
}

template <class STATE>
void StateFunction::Destroy(STATE &state, duckdb::AggregateInputData &) {
  if (state.arena) state.arena->Release();
  state.arena = nullptr;
}
template void StateFunction::Destroy<SumState>(SumState &, duckdb::AggregateInputData &);

duckdb::unique_ptr<duckdb::FunctionData> SumNoLiftBind(duckdb::ClientContext &, duckdb::AggregateFunction &function,
                                                       duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &) {
  function.return_type = ResultType(false);
  return duckdb::make_uniq<duckdb::VariableReturnBindData>(function.return_type);
}

duckdb::unique_ptr<duckdb::FunctionData> sum_to_nb_agg_bind(duckdb::ClientContext &, duckdb::AggregateFunction &function,
                                                            duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &) {
  function.return_type = ResultType(true);
  return duckdb::make_uniq<duckdb::VariableReturnBindData>(function.return_type);
}

void SumNoLift(duckdb::Vector inputs[], duckdb::AggregateInputData &, idx_t input_count, duckdb::Vector &state_vector,
               idx_t count) {
  Update(CFB_TRIPLE, inputs, input_count, state_vector, count);
}

void sum_to_nb_agg(duckdb::Vector inputs[], duckdb::AggregateInputData &, duckdb::idx_t cols, duckdb::Vector &state_vector,
                   duckdb::idx_t count) {
  Update(CFB_NB, inputs, cols, state_vector, count);
}

void SumNoLiftSimple(duckdb::Vector inputs[], duckdb::AggregateInputData &, idx_t input_count, duckdb::data_ptr_t state,
                     idx_t count) {
  SimpleUpdate(CFB_TRIPLE, inputs, input_count, state, count);
}
void sum_to_nb_agg_simple(duckdb::Vector inputs[], duckdb::AggregateInputData &, idx_t input_count, duckdb::data_ptr_t state,
                          idx_t count) {
  SimpleUpdate(CFB_NB, inputs, input_count, state, count);
}

void SumStateCombine(duckdb::Vector &state, duckdb::Vector &combined, duckdb::AggregateInputData &, idx_t count) {
  duckdb::UnifiedVectorFormat sdata;
  state.ToUnifiedFormat(count, sdata);
  auto src = (SumState **)sdata.data;
  auto dst = duckdb::FlatVector::GetData<SumState *>(combined);
  // states that live in the same pair of arenas are merged with one call (slot lists)
  struct Batch {
    Arena *d, *s;
    std::vector<int32_t> ds, ss;
  };
  std::vector<Batch> batches;
  for (idx_t i = 0; i < count; i++) {
    SumState *s = src[sdata.sel->get_index(i)];
    if (!s->arena) continue;  // the source never saw a row
    if (!dst[i]->arena) {
      // empty target adopts the source's slot (sum_state.cpp:26-60); the source stays destroyable:
      // its destructor sees a null handle
      dst[i]->arena = s->arena;
      dst[i]->slot = s->slot;
      s->arena = nullptr;
      continue;
    }
    size_t b = 0;
    while (b < batches.size() && (batches[b].d != dst[i]->arena || batches[b].s != s->arena)) b++;
    if (b == batches.size()) batches.push_back(Batch{dst[i]->arena, s->arena, {}, {}});
    batches[b].ds.push_back(dst[i]->slot);
    batches[b].ss.push_back(s->slot);
  }
  for (auto &b : batches) {
    if (b.d == b.s) {
      // two states of one arena (a group the hash aggregate emitted twice from one thread): merged inside
      // the context, one pair per call -- a slot may be the target of one pair and the source of the next
      std::lock_guard<std::mutex> g(b.d->mu);
      for (size_t i = 0; i < b.ds.size(); i++) Check(cfb_ctx_combine_slots(b.d->ctx, b.d->ctx, 1, &b.ds[i], &b.ss[i]));
      continue;
    }
    // both contexts are entered: lock both arenas, in address order
    std::mutex &m1 = b.d < b.s ? b.d->mu : b.s->mu, &m2 = b.d < b.s ? b.s->mu : b.d->mu;
    std::lock_guard<std::mutex> g1(m1), g2(m2);
    Check(cfb_ctx_combine_slots(b.d->ctx, b.s->ctx, b.ds.size(), b.ds.data(), b.ss.data()));
  }
}

// Writes `count` canonical results into the nested STRUCT vector in the reference's layout
// (sum_state.cpp:132-461); shared by finalize and by multiply_triple / multiply_nb_agg.
void WriteResults(const std::vector<cfb_result> &res, duckdb::Vector &result, bool nb, int n, int m) {
  using namespace duckdb;
  const idx_t count = res.size();
  idx_t total_keys = 0, total_pairs = 0;
  for (const cfb_result &r : res) {
    total_keys += (idx_t)r.total_keys;
    if (r.pair_offsets) total_pairs += (idx_t)r.pair_offsets[r.n_pair_lists];
  }
  const idx_t nq = nb ? (idx_t)n : (idx_t)n * (n + 1) / 2;
  const idx_t npl = (idx_t)m * (m + 1) / 2;

  auto &kids = StructVector::GetEntries(result);
  Vector &vN = *kids[0], &vLin = *kids[1], &vQuad = *kids[2], &vLinCat = *kids[3];
  auto N = FlatVector::GetData<int32_t>(vN);

  ListVector::Reserve(vLin, (idx_t)n * count);
  ListVector::SetListSize(vLin, (idx_t)n * count);
  ListVector::Reserve(vQuad, nq * count);
  ListVector::SetListSize(vQuad, nq * count);
  auto lin_e = ListVector::GetData(vLin);
  auto quad_e = ListVector::GetData(vQuad);
  auto lin_d = FlatVector::GetData<float>(ListVector::GetEntry(vLin));
  auto quad_d = FlatVector::GetData<float>(ListVector::GetEntry(vQuad));

  // lin_cat: LIST(LIST(STRUCT(key,value))): outer = m lists per group, inner = keys
  ListVector::Reserve(vLinCat, (idx_t)m * count);
  ListVector::SetListSize(vLinCat, (idx_t)m * count);
  Vector &lc_inner = ListVector::GetEntry(vLinCat);
  ListVector::Reserve(lc_inner, total_keys);
  ListVector::SetListSize(lc_inner, total_keys);
  auto lc_outer_e = ListVector::GetData(vLinCat);
  auto lc_inner_e = ListVector::GetData(lc_inner);
  auto &lc_kv = StructVector::GetEntries(ListVector::GetEntry(lc_inner));
  auto lc_key = FlatVector::GetData<int32_t>(*lc_kv[0]);
  auto lc_val = FlatVector::GetData<float>(*lc_kv[1]);

  list_entry_t *nc_outer_e = nullptr, *nc_inner_e = nullptr, *cc_outer_e = nullptr, *cc_inner_e = nullptr;
  int32_t *nc_key = nullptr, *cc_k1 = nullptr, *cc_k2 = nullptr;
  float *nc_val = nullptr, *cc_val = nullptr;
  if (!nb) {
    Vector &vNumCat = *kids[4], &vCatCat = *kids[5];
    ListVector::Reserve(vNumCat, (idx_t)n * m * count);
    ListVector::SetListSize(vNumCat, (idx_t)n * m * count);
    Vector &nc_inner = ListVector::GetEntry(vNumCat);
    ListVector::Reserve(nc_inner, total_keys * n);
    ListVector::SetListSize(nc_inner, total_keys * n);
    nc_outer_e = ListVector::GetData(vNumCat);
    nc_inner_e = ListVector::GetData(nc_inner);
    auto &nc_kv = StructVector::GetEntries(ListVector::GetEntry(nc_inner));
    nc_key = FlatVector::GetData<int32_t>(*nc_kv[0]);
    nc_val = FlatVector::GetData<float>(*nc_kv[1]);
    ListVector::Reserve(vCatCat, npl * count);
    ListVector::SetListSize(vCatCat, npl * count);
    Vector &cc_inner = ListVector::GetEntry(vCatCat);
    ListVector::Reserve(cc_inner, total_pairs);
    ListVector::SetListSize(cc_inner, total_pairs);
    cc_outer_e = ListVector::GetData(vCatCat);
    cc_inner_e = ListVector::GetData(cc_inner);
    auto &cc_kkv = StructVector::GetEntries(ListVector::GetEntry(cc_inner));
    cc_k1 = FlatVector::GetData<int32_t>(*cc_kkv[0]);
    cc_k2 = FlatVector::GetData<int32_t>(*cc_kkv[1]);
    cc_val = FlatVector::GetData<float>(*cc_kkv[2]);
  }

  idx_t key_pos = 0, nc_pos = 0, cc_pos = 0;
  for (idx_t i = 0; i < count; i++) {
    const cfb_result &r = res[i];
    const bool live = r.lin != nullptr;
    N[i] = (int32_t)r.N;  // INTEGER in the STRUCT (sum_no_lift.cpp:22)
    lin_e[i] = {i * (idx_t)n, (idx_t)n};
    quad_e[i] = {i * nq, nq};
    for (int k = 0; k < n; k++) lin_d[i * n + k] = live ? (float)r.lin[k] : 0.f;
    for (idx_t k = 0; k < nq; k++) quad_d[i * nq + k] = live ? (float)r.quad[k] : 0.f;
    // lin_cat[c] = [{key, count}], keys ascending (sum_state.cpp:372-395)
    lc_outer_e[i] = {i * (idx_t)m, (idx_t)m};
    const idx_t group_key0 = key_pos;
    for (int c = 0; c < m; c++) {
      const idx_t lo = live ? (idx_t)r.cat_offsets[c] : 0, hi = live ? (idx_t)r.cat_offsets[c + 1] : 0;
      lc_inner_e[i * m + c] = {key_pos, hi - lo};
      for (idx_t t = lo; t < hi; t++, key_pos++) {
        lc_key[key_pos] = r.cat_keys[t];
        lc_val[key_pos] = (float)r.cat_counts[t];
      }
    }
    (void)group_key0;
    if (nb) continue;
    // quad_num_cat: sub-list index num*m + cat (sum_state.cpp:383-404); value = sum x_num | key
    nc_outer_e[i] = {i * (idx_t)n * m, (idx_t)n * m};
    for (int k = 0; k < n; k++)
      for (int c = 0; c < m; c++) {
        const idx_t lo = live ? (idx_t)r.cat_offsets[c] : 0, hi = live ? (idx_t)r.cat_offsets[c + 1] : 0;
        nc_inner_e[(i * n + k) * m + c] = {nc_pos, hi - lo};
        for (idx_t t = lo; t < hi; t++, nc_pos++) {
          nc_key[nc_pos] = r.cat_keys[t];
          nc_val[nc_pos] = (float)r.numcat_sums[(idx_t)k * r.total_keys + t];
        }
      }
    // quad_cat: m(m+1)/2 lists in (k<=l) order, entries ascending by (key1,key2) (sum_state.cpp:440-461)
    cc_outer_e[i] = {i * npl, npl};
    for (idx_t p = 0; p < npl; p++) {
      const idx_t lo = live ? (idx_t)r.pair_offsets[p] : 0, hi = live ? (idx_t)r.pair_offsets[p + 1] : 0;
      cc_inner_e[i * npl + p] = {cc_pos, hi - lo};
      for (idx_t t = lo; t < hi; t++, cc_pos++) {
        cc_k1[cc_pos] = r.pair_key1[t];
        cc_k2[cc_pos] = r.pair_key2[t];
        cc_val[cc_pos] = (float)r.pair_counts[t];
      }
    }
  }
}

void SumStateFinalize(duckdb::Vector &state_vector, duckdb::AggregateInputData &, duckdb::Vector &result, idx_t count,
                      idx_t offset) {
  using namespace duckdb;
  if (offset != 0) throw InternalException("ring aggregate finalize expects offset 0");  // sum_state.cpp:120
  if (count == 0) return;
  UnifiedVectorFormat sdata;
  state_vector.ToUnifiedFormat(count, sdata);
  auto states = (SumState **)sdata.data;

  // Pull every group's canonical result first: list children are sized once, then filled.
  std::vector<cfb_result> res(count);
  struct Guard {
    std::vector<cfb_result> &r;
    ~Guard() {
      for (auto &x : r) cfb_result_free(&x);
    }
  } guard{res};
  const bool nb = StructVector::GetEntries(result).size() == 4;
  int n = 0, m = 0;
  for (idx_t i = 0; i < count; i++) {
    SumState *s = states[sdata.sel->get_index(i)];
    memset(&res[i], 0, sizeof(cfb_result));
    if (s->arena) {
      std::lock_guard<std::mutex> g(s->arena->mu);
      Check(cfb_ctx_finalize(s->arena->ctx, s->slot, &res[i]));
      n = res[i].n_num;
      m = res[i].n_cat;
    }
  }
  WriteResults(res, result, nb, n, m);
}

}  // namespace Triple
