// replay_host.cpp -- see replay_host.h.
#include "replay_host.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>

#include <duckdb.hpp>

// Provided by the extension linked into this library (ours or the reference's).
namespace duckdb_ring {
void Load(duckdb::DatabaseInstance &db);
const char *Implementation();
}  // namespace duckdb_ring
// The loadable-extension entry points (our build exports them; the reference's test build does not).
extern "C" __attribute__((weak)) void duckdb_imputation_init(duckdb::DatabaseInstance &db);
extern "C" __attribute__((weak)) const char *duckdb_imputation_version();

namespace {

thread_local std::string g_error;
duckdb::DatabaseInstance &Catalog() {
  static duckdb::DatabaseInstance db;
  static std::once_flag once;
  std::call_once(once, [] { duckdb_ring::Load(db); });
  return db;
}

void RenderValue(duckdb::Vector &v, idx_t row, std::ostringstream &os) {
  using namespace duckdb;
  const auto &t = v.GetType();
  switch (t.id()) {
    case LogicalTypeId::INTEGER:
      os << FlatVector::GetData<int32_t>(v)[row];
      break;
    case LogicalTypeId::FLOAT: {
      char buf[64];
      snprintf(buf, sizeof(buf), "%.9g", (double)FlatVector::GetData<float>(v)[row]);
      os << buf;
      break;
    }
    case LogicalTypeId::LIST: {
      const list_entry_t e = ListVector::GetData(v)[row];
      Vector &child = ListVector::GetEntry(v);
      os << "[";
      for (idx_t i = 0; i < e.length; i++) {
        if (i) os << ", ";
        RenderValue(child, e.offset + i, os);
      }
      os << "]";
      break;
    }
    case LogicalTypeId::STRUCT: {
      auto &kids = StructVector::GetEntries(v);
      os << "{";
      for (size_t i = 0; i < kids.size(); i++) {
        if (i) os << ", ";
        os << "\"" << t.children()[i].first << "\": ";
        RenderValue(*kids[i], row, os);
      }
      os << "}";
      break;
    }
    default:
      throw duckdb::InternalException("replay: cannot render this type");
  }
}

struct ThreadLocalStates {
  std::unique_ptr<duckdb::data_t[]> mem;  // [split][n_groups] * state_size, like hash-table row storage
  std::vector<char> live;
};

// Shapes of DuckDB's protocol that a plain scan does not produce (replay_set_option).
struct Options {
  int no_simple = 0;          // 1: never use simple_update (force the hash-aggregate protocol)
  int lift_shape = 0;         // lifted argument: 0 flat, 1 DICTIONARY struct (reversed rows), 2 flat struct with
                              // DICTIONARY children and leaves, 3 CONSTANT struct (a CROSS JOIN side)
  int split_states = 1;       // k > 1: a worker keeps k state sets per group (hash table reset when full) and
                              // merges them itself: two states of one group from one thread
  int parallel_finalize = 0;  // P > 1: P threads combine + finalize disjoint groups at once (radix partitions)
};
Options g_opt;

}  // namespace

extern "C" {

const char *replay_last_error(void) { return g_error.c_str(); }

int replay_set_option(const char *name, int value) {
  const std::string n = name ? name : "";
  if (n == "no_simple") g_opt.no_simple = value;
  else if (n == "lift_shape") g_opt.lift_shape = value;
  else if (n == "split_states") g_opt.split_states = value < 1 ? 1 : value;
  else if (n == "parallel_finalize") g_opt.parallel_finalize = value;
  else if (n == "reset") g_opt = Options{};
  else {
    g_error = "replay: unknown option " + n;
    return -1;
  }
  return 0;
}

// Load a fresh catalog through the extension's exported entry points; returns the number of functions registered
// (-1: the entry points are not exported by this build).  version_out: duckdb_imputation_version().
int replay_load_via_entry_points(const char **version_out) {
  if (!duckdb_imputation_init || !duckdb_imputation_version) {
    g_error = "this build does not export duckdb_imputation_init / duckdb_imputation_version";
    return -1;
  }
  duckdb::DatabaseInstance db;
  duckdb_imputation_init(db);
  if (version_out) *version_out = duckdb_imputation_version();
  return (int)(db.aggregates.size() + db.scalars.size());
}
const char *replay_implementation(void) { return duckdb_ring::Implementation(); }
void replay_free(char *p) { free(p); }

char *replay_list_functions(void) {
  std::string s;
  for (auto &kv : Catalog().aggregates) s += kv.first + "\n";
  return strdup(s.c_str());
}

namespace {
// bind + evaluate a registered scalar function on one chunk
struct BoundScalar {
  duckdb::ScalarFunction fun;
  duckdb::unique_ptr<duckdb::FunctionData> bind_data;
};
std::unique_ptr<BoundScalar> BindScalar(const char *name, int n_num, int n_cat) {
  using namespace duckdb;
  auto it = Catalog().scalars.find(name);
  if (it == Catalog().scalars.end())
    throw InvalidInputException(std::string("Catalog Error: scalar function ") + name + " does not exist");
  std::unique_ptr<BoundScalar> bp(new BoundScalar{it->second, nullptr});
  BoundScalar &b = *bp;
  ClientContext context;
  vector<unique_ptr<Expression>> args;
  for (int k = 0; k < n_num + n_cat; k++) {
    args.push_back(make_uniq<Expression>());
    args.back()->return_type = k < n_num ? LogicalType::FLOAT : LogicalType::INTEGER;
  }
  if (b.fun.bind) b.bind_data = b.fun.bind(context, b.fun, args);
  return bp;
}
void FillChunk(duckdb::DataChunk &chunk, int n_num, int n_cat, const float *const *num, const int32_t *const *cat,
               size_t lo, duckdb::sel_t *rel, idx_t count) {
  using namespace duckdb;
  chunk.data.clear();
  for (int k = 0; k < n_num; k++) chunk.data.emplace_back(LogicalType::FLOAT, (data_ptr_t)(num[k] + lo));
  for (int k = 0; k < n_cat; k++) chunk.data.emplace_back(LogicalType::INTEGER, (data_ptr_t)(cat[k] + lo));
  if (rel)
    for (auto &v : chunk.data) v.Slice(SelectionVector(rel));
  chunk.SetCardinality(count);
}
}  // namespace

int replay_scalar(const char *scalar, int n_num, int n_cat, const float *const *num, const int32_t *const *cat,
                  const uint32_t *sel, size_t n_sel, size_t rows, char **json_out) {
  using namespace duckdb;
  try {
    if (!scalar || !json_out) throw InvalidInputException("bad arguments");
    *json_out = nullptr;
    auto bp = BindScalar(scalar, n_num, n_cat);
    BoundScalar &b = *bp;
    std::ostringstream os;
    os << "[";
    bool first = true;
    std::vector<sel_t> rel(STANDARD_VECTOR_SIZE);
    size_t s_pos = 0;
    for (size_t lo = 0; lo < rows; lo += STANDARD_VECTOR_SIZE) {
      const size_t hi = std::min(rows, lo + STANDARD_VECTOR_SIZE);
      idx_t count = hi - lo;
      if (sel) {
        count = 0;
        while (s_pos < n_sel && sel[s_pos] < hi) rel[count++] = (sel_t)(sel[s_pos++] - lo);
        if (!count) continue;
      }
      DataChunk chunk;
      FillChunk(chunk, n_num, n_cat, num, cat, lo, sel ? rel.data() : nullptr, count);
      ExpressionState state;
      Vector result(b.fun.return_type, count);
      b.fun.function(chunk, state, result);
      for (idx_t r = 0; r < count; r++) {
        if (!first) os << ", ";
        first = false;
        RenderValue(result, r, os);
      }
    }
    os << "]";
    *json_out = strdup(os.str().c_str());
    return 0;
  } catch (std::exception &e) {
    g_error = e.what();
    return -1;
  }
}

namespace {
// Type-directed reader of the JSON RenderValue writes: fills row `row` of `v`.
void SkipWs(const char *&p) {
  while (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r') p++;
}
void Expect(const char *&p, char c) {
  SkipWs(p);
  if (*p != c) throw duckdb::InvalidInputException(std::string("replay: malformed STRUCT literal, expected '") + c + "'");
  p++;
}
void ParseValue(const char *&p, duckdb::Vector &v, idx_t row) {
  using namespace duckdb;
  SkipWs(p);
  switch (v.GetType().id()) {
    case LogicalTypeId::INTEGER: {
      char *end;
      FlatVector::GetData<int32_t>(v)[row] = (int32_t)strtol(p, &end, 10);
      p = end;
      break;
    }
    case LogicalTypeId::FLOAT: {
      char *end;
      FlatVector::GetData<float>(v)[row] = (float)strtod(p, &end);
      p = end;
      break;
    }
    case LogicalTypeId::LIST: {
      Expect(p, '[');
      const idx_t start = ListVector::GetListSize(v);
      idx_t len = 0;
      SkipWs(p);
      while (*p != ']') {
        if (len) Expect(p, ',');
        ListVector::Reserve(v, start + len + 1);
        ListVector::SetListSize(v, start + len + 1);
        ParseValue(p, ListVector::GetEntry(v), start + len);
        len++;
        SkipWs(p);
      }
      p++;
      ListVector::GetData(v)[row] = {start, len};
      break;
    }
    case LogicalTypeId::STRUCT: {
      Expect(p, '{');
      auto &kids = StructVector::GetEntries(v);
      for (size_t i = 0; i < kids.size(); i++) {
        if (i) Expect(p, ',');
        Expect(p, '"');
        while (*p && *p != '"') p++;  // field names are not checked: the ring functions read children positionally
        Expect(p, '"');
        Expect(p, ':');
        ParseValue(p, *kids[i], row);
      }
      Expect(p, '}');
      break;
    }
    default:
      throw InternalException("replay: cannot parse this type");
  }
}
duckdb::LogicalType RingType(bool nb) {
  using namespace duckdb;
  child_list_t<LogicalType> kv;
  kv.emplace_back("key", LogicalType::INTEGER);
  kv.emplace_back("value", LogicalType::FLOAT);
  child_list_t<LogicalType> f;
  f.emplace_back("N", LogicalType::INTEGER);
  f.emplace_back("lin_agg", LogicalType::LIST(LogicalType::FLOAT));
  f.emplace_back("quad_agg", LogicalType::LIST(LogicalType::FLOAT));
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

int replay_predict(const char *scalar, const float *params, size_t n_params, const int *flags, int n_flags, int n_num,
                   int n_cat, const float *const *num, const int32_t *const *cat, const uint32_t *sel, size_t n_sel,
                   size_t rows, void *out) {
  using namespace duckdb;
  try {
    if (!scalar || !out || (!params && n_params)) throw InvalidInputException("bad arguments");
    auto it = Catalog().scalars.find(scalar);
    if (it == Catalog().scalars.end())
      throw InvalidInputException(std::string("Catalog Error: scalar function ") + scalar + " does not exist");
    ScalarFunction fun = it->second;
    ClientContext context;
    vector<unique_ptr<Expression>> args;
    unique_ptr<FunctionData> bind_data;
    if (fun.bind) bind_data = fun.bind(context, fun, args);
    const idx_t out_w = fun.return_type.width();
    ExpressionState state;  // one executor: its local state lives as long as the query
    if (fun.init_local_state) state.local = fun.init_local_state(state, BoundFunctionExpression(), bind_data.get());
    // the literal arguments: a constant FLOAT[] and constant BOOLEANs, as the planner hands them over
    Vector plist(LogicalType::LIST(LogicalType::FLOAT), 1);
    ListVector::Reserve(plist, n_params);
    ListVector::SetListSize(plist, n_params);
    if (n_params) memcpy(FlatVector::GetData<float>(ListVector::GetEntry(plist)), params, n_params * sizeof(float));
    ListVector::GetData(plist)[0] = {0, n_params};
    plist.SetVectorType(VectorType::CONSTANT_VECTOR);
    std::vector<Vector> lits;
    for (int i = 0; i < n_flags; i++) {
      lits.emplace_back(LogicalType::BOOLEAN, 1);
      FlatVector::GetData<uint8_t>(lits.back())[0] = flags[i] ? 1 : 0;
      lits.back().SetVectorType(VectorType::CONSTANT_VECTOR);
    }
    std::vector<sel_t> rel(STANDARD_VECTOR_SIZE);
    size_t s_pos = 0, written = 0;
    for (size_t lo = 0; lo < rows; lo += STANDARD_VECTOR_SIZE) {
      const size_t hi = std::min(rows, lo + STANDARD_VECTOR_SIZE);
      idx_t count = hi - lo;
      if (sel) {
        count = 0;
        while (s_pos < n_sel && sel[s_pos] < hi) rel[count++] = (sel_t)(sel[s_pos++] - lo);
        if (!count) continue;
      }
      DataChunk cols;
      FillChunk(cols, n_num, n_cat, num, cat, lo, sel ? rel.data() : nullptr, count);
      DataChunk chunk;
      chunk.data.push_back(plist);
      for (auto &l : lits) chunk.data.push_back(l);
      for (auto &v : cols.data) chunk.data.push_back(v);
      chunk.SetCardinality(count);
      Vector result(fun.return_type, count);
      fun.function(chunk, state, result);
      memcpy((char *)out + written * out_w, FlatVector::GetData(result), count * out_w);
      written += count;
    }
    return 0;
  } catch (std::exception &e) {
    g_error = e.what();
    return -1;
  }
}

int replay_train(const char *scalar, const char *json_triple, int n_consts, const double *consts, const char *const_types,
                 float **params_out, size_t *n_params_out) {
  using namespace duckdb;
  try {
    if (!scalar || !json_triple || !params_out || !n_params_out || (n_consts && (!consts || !const_types)))
      throw InvalidInputException("bad arguments");
    *params_out = nullptr;
    *n_params_out = 0;
    auto it = Catalog().scalars.find(scalar);
    if (it == Catalog().scalars.end())
      throw InvalidInputException(std::string("Catalog Error: scalar function ") + scalar + " does not exist");
    ScalarFunction fun = it->second;
    const LogicalType arg_type = RingType(false);
    ClientContext context;
    vector<unique_ptr<Expression>> args;
    args.push_back(make_uniq<Expression>());
    args.back()->return_type = arg_type;
    unique_ptr<FunctionData> bind_data;
    if (fun.bind) bind_data = fun.bind(context, fun, args);
    // SELECT <scalar>(<STRUCT literal>, <constants>): one row, every argument a CONSTANT vector but the STRUCT
    DataChunk chunk;
    chunk.data.emplace_back(arg_type, 1);
    const char *p = json_triple;
    ParseValue(p, chunk.data[0], 0);
    for (int i = 0; i < n_consts; i++) {
      switch (const_types[i]) {
        case 'i':
          chunk.data.emplace_back(LogicalType::INTEGER, 1);
          FlatVector::GetData<int32_t>(chunk.data.back())[0] = (int32_t)consts[i];
          break;
        case 'f':
          chunk.data.emplace_back(LogicalType::FLOAT, 1);
          FlatVector::GetData<float>(chunk.data.back())[0] = (float)consts[i];
          break;
        case 'b':
          chunk.data.emplace_back(LogicalType::BOOLEAN, 1);
          FlatVector::GetData<uint8_t>(chunk.data.back())[0] = consts[i] != 0.0;
          break;
        default:
          throw InvalidInputException("replay_train: constant types are 'i', 'f' or 'b'");
      }
      chunk.data.back().SetVectorType(VectorType::CONSTANT_VECTOR);
    }
    chunk.SetCardinality(1);
    ExpressionState state;
    Vector result(fun.return_type.id() == LogicalTypeId::LIST && !fun.return_type.children().empty()
                      ? fun.return_type
                      : LogicalType::LIST(LogicalType::FLOAT),
                  1);
    fun.function(chunk, state, result);
    const list_entry_t e = ListVector::GetData(result)[0];
    float *out = (float *)malloc(std::max<size_t>(1, e.length) * sizeof(float));
    memcpy(out, FlatVector::GetData<float>(ListVector::GetEntry(result)) + e.offset, e.length * sizeof(float));
    *params_out = out;
    *n_params_out = e.length;
    return 0;
  } catch (std::exception &e) {
    g_error = e.what();
    return -1;
  }
}

int replay_train_list(const char *scalar, int nb, const char *json_triples, const int32_t *labels, size_t n_labels,
                      int n_consts, const double *consts, const char *const_types, float **params_out, size_t *n_params_out) {
  using namespace duckdb;
  try {
    if (!scalar || !json_triples || !params_out || !n_params_out || (n_labels && !labels) || (n_consts && (!consts || !const_types)))
      throw InvalidInputException("bad arguments");
    *params_out = nullptr;
    *n_params_out = 0;
    auto it = Catalog().scalars.find(scalar);
    if (it == Catalog().scalars.end())
      throw InvalidInputException(std::string("Catalog Error: scalar function ") + scalar + " does not exist");
    ScalarFunction fun = it->second;
    const LogicalType list_type = LogicalType::LIST(RingType(nb != 0));
    ClientContext context;
    vector<unique_ptr<Expression>> args;
    args.push_back(make_uniq<Expression>());
    args.back()->return_type = list_type;
    unique_ptr<FunctionData> bind_data;
    if (fun.bind) bind_data = fun.bind(context, fun, args);
    // SELECT <scalar>(list(agg), list(label), <constants>): one row
    DataChunk chunk;
    chunk.data.emplace_back(list_type, 1);
    const char *p = json_triples;
    ParseValue(p, chunk.data[0], 0);
    chunk.data.emplace_back(LogicalType::LIST(LogicalType::INTEGER), 1);
    ListVector::Reserve(chunk.data[1], n_labels);
    ListVector::SetListSize(chunk.data[1], n_labels);
    if (n_labels) memcpy(FlatVector::GetData<int32_t>(ListVector::GetEntry(chunk.data[1])), labels, n_labels * sizeof(int32_t));
    ListVector::GetData(chunk.data[1])[0] = {0, n_labels};
    for (int i = 0; i < n_consts; i++) {
      switch (const_types[i]) {
        case 'i':
          chunk.data.emplace_back(LogicalType::INTEGER, 1);
          FlatVector::GetData<int32_t>(chunk.data.back())[0] = (int32_t)consts[i];
          break;
        case 'f':
          chunk.data.emplace_back(LogicalType::FLOAT, 1);
          FlatVector::GetData<float>(chunk.data.back())[0] = (float)consts[i];
          break;
        case 'b':
          chunk.data.emplace_back(LogicalType::BOOLEAN, 1);
          FlatVector::GetData<uint8_t>(chunk.data.back())[0] = consts[i] != 0.0;
          break;
        default:
          throw InvalidInputException("replay_train_list: constant types are 'i', 'f' or 'b'");
      }
      chunk.data.back().SetVectorType(VectorType::CONSTANT_VECTOR);
    }
    chunk.SetCardinality(1);
    ExpressionState state;
    Vector result(LogicalType::LIST(LogicalType::FLOAT), 1);
    fun.function(chunk, state, result);
    const list_entry_t e = ListVector::GetData(result)[0];
    float *out = (float *)malloc(std::max<size_t>(1, e.length) * sizeof(float));
    memcpy(out, FlatVector::GetData<float>(ListVector::GetEntry(result)) + e.offset, e.length * sizeof(float));
    *params_out = out;
    *n_params_out = e.length;
    return 0;
  } catch (std::exception &e) {
    g_error = e.what();
    return -1;
  }
}

int replay_scalar_structs(const char *scalar, int nb, int n_args, const char *const *json_args, size_t rows, char **json_out) {
  using namespace duckdb;
  try {
    if (!scalar || !json_out || n_args < 1 || !json_args) throw InvalidInputException("bad arguments");
    *json_out = nullptr;
    auto it = Catalog().scalars.find(scalar);
    if (it == Catalog().scalars.end())
      throw InvalidInputException(std::string("Catalog Error: scalar function ") + scalar + " does not exist");
    ScalarFunction fun = it->second;
    const LogicalType arg_type = RingType(nb != 0);
    ClientContext context;
    vector<unique_ptr<Expression>> args;
    for (int k = 0; k < n_args; k++) {
      args.push_back(make_uniq<Expression>());
      args.back()->return_type = arg_type;
    }
    unique_ptr<FunctionData> bind_data;
    if (fun.bind) bind_data = fun.bind(context, fun, args);
    std::ostringstream os;
    os << "[";
    std::vector<const char *> cur(json_args, json_args + n_args);
    for (auto &p : cur) Expect(p, '[');
    for (size_t lo = 0; lo < rows; lo += STANDARD_VECTOR_SIZE) {
      const idx_t count = std::min<size_t>(rows - lo, STANDARD_VECTOR_SIZE);
      DataChunk chunk;
      for (int k = 0; k < n_args; k++) {
        chunk.data.emplace_back(arg_type, count);
        for (idx_t r = 0; r < count; r++) {
          if (lo + r) Expect(cur[k], ',');
          ParseValue(cur[k], chunk.data[k], r);
        }
      }
      chunk.SetCardinality(count);
      // what a join hands a scalar function (replay_set_option lift_shape): every argument behind a selection
      // that reverses the rows (1: on the STRUCT, 2: on its children), or the LAST argument as a CONSTANT vector --
      // its first row for every row of the chunk (3: a CROSS JOIN side).  Output row r then belongs to input row
      // count-1-r (1, 2); the rendering below puts it back in input order.
      std::vector<sel_t> rev(count);
      for (idx_t r = 0; r < count; r++) rev[r] = (sel_t)(count - 1 - r);
      const int shape = g_opt.lift_shape;
      if (shape == 1)
        for (auto &v : chunk.data) v.Slice(SelectionVector(rev.data()));
      else if (shape == 2)
        for (auto &v : chunk.data)
          for (auto &kid : StructVector::GetEntries(v)) kid->Slice(SelectionVector(rev.data()));
      else if (shape == 3)
        chunk.data.back().SetVectorType(VectorType::CONSTANT_VECTOR);
      ExpressionState state;
      Vector result(fun.return_type, count);
      fun.function(chunk, state, result);
      for (idx_t r = 0; r < count; r++) {
        if (lo + r) os << ", ";
        RenderValue(result, (shape == 1 || shape == 2) ? count - 1 - r : r, os);
      }
    }
    os << "]";
    *json_out = strdup(os.str().c_str());
    return 0;
  } catch (std::exception &e) {
    g_error = e.what();
    return -1;
  }
}

int replay_aggregate(const char *function, const char *scalar, int n_num, int n_cat, const float *const *num,
                     const int32_t *const *cat, const int32_t *group, int n_groups, const uint32_t *sel,
                     size_t n_sel, size_t rows, int threads, char **json_out, double *seconds) {
  using namespace duckdb;
  try {
    if (!function || !json_out || n_groups < 1) throw InvalidInputException("bad arguments");
    *json_out = nullptr;
    auto it = Catalog().aggregates.find(function);
    if (it == Catalog().aggregates.end())
      throw InvalidInputException(std::string("Catalog Error: aggregate function ") + function + " does not exist");
    AggregateFunction fun = it->second;  // bind mutates its copy (return_type)
    const int n_cols = n_num + n_cat;
    const bool lifted = scalar && *scalar;
    if (!lifted && (size_t)n_cols != fun.arguments.size() && fun.varargs.id() == LogicalTypeId::INVALID)
      throw InvalidInputException("argument count does not match the function signature");
    std::unique_ptr<BoundScalar> lift;
    if (lifted) lift = BindScalar(scalar, n_num, n_cat);
    ClientContext context;
    vector<unique_ptr<Expression>> args;
    unique_ptr<FunctionData> bind_data;
    if (fun.bind) bind_data = fun.bind(context, fun, args);
    AggregateInputData aggr(bind_data.get());
    const idx_t ssz = fun.state_size();

    const size_t n_chunks = (rows + STANDARD_VECTOR_SIZE - 1) / STANDARD_VECTOR_SIZE;
    int T = std::max(1, threads);
    if ((size_t)T > std::max<size_t>(1, n_chunks)) T = (int)std::max<size_t>(1, n_chunks);
    const Options opt = g_opt;
    const int K = opt.split_states;  // state sets per worker
    // a query without GROUP BY over an aggregate with simple_update is planned as an ungrouped aggregate
    const bool simple = !group && n_groups == 1 && fun.simple_update && !opt.no_simple && K == 1;
    std::vector<ThreadLocalStates> tls(T);
    for (auto &t : tls) {
      t.mem.reset(new data_t[std::max<size_t>(1, ssz * n_groups * K)]);
      t.live.assign((size_t)n_groups * K, 0);
    }
    std::string worker_error;
    std::mutex err_mu;

    const auto t0 = std::chrono::steady_clock::now();
    auto worker = [&](int t) {
      try {
        ThreadLocalStates &loc = tls[t];
        const size_t c_lo = n_chunks * t / T, c_hi = n_chunks * (t + 1) / T;
        std::vector<data_ptr_t> ptrs(STANDARD_VECTOR_SIZE), ptrs2(STANDARD_VECTOR_SIZE);
        std::vector<sel_t> rel(STANDARD_VECTOR_SIZE), rev(STANDARD_VECTOR_SIZE), ident(4 * STANDARD_VECTOR_SIZE);
        for (size_t i = 0; i < ident.size(); i++) ident[i] = (sel_t)i;
        // first selected row at or after this thread's range (sel is ascending)
        size_t s_pos = sel ? (size_t)(std::lower_bound(sel, sel + n_sel, (uint32_t)(c_lo * STANDARD_VECTOR_SIZE)) - sel) : 0;
        for (size_t c = c_lo; c < c_hi; c++) {
          const size_t lo = c * STANDARD_VECTOR_SIZE, hi = std::min(rows, lo + STANDARD_VECTOR_SIZE);
          idx_t count = hi - lo;
          if (sel) {
            count = 0;
            while (s_pos < n_sel && sel[s_pos] < hi) rel[count++] = (sel_t)(sel[s_pos++] - lo);
            if (!count) continue;
          }
          const size_t set = K > 1 ? c % (size_t)K : 0;
          for (idx_t r = 0; r < count; r++) {
            const size_t row = lo + (sel ? rel[r] : r);
            const int g = group ? group[row] : 0;
            if (g < 0 || g >= n_groups) throw InvalidInputException("group slot out of range");
            data_ptr_t st = loc.mem.get() + (set * n_groups + (size_t)g) * ssz;
            if (!loc.live[set * n_groups + g]) {
              fun.initialize(st);
              loc.live[set * n_groups + g] = 1;
            }
            ptrs[r] = st;
          }
          // scan vectors over the table's column storage (+ the filter's selection on top)
          DataChunk chunk;
          FillChunk(chunk, n_num, n_cat, num, cat, lo, sel ? rel.data() : nullptr, count);
          Vector state_vector(LogicalType::POINTER, (data_ptr_t)ptrs.data());
          if (lifted) {
            ExpressionState estate;
            std::vector<Vector> one;
            one.emplace_back(lift->fun.return_type, count);
            lift->fun.function(chunk, estate, one[0]);
            if (opt.lift_shape == 1 || opt.lift_shape == 2) {
              // rows reach the aggregate in reverse order through a selection: on the STRUCT itself (1) or on each
              // of its children, with an explicit identity selection on every leaf below them (2)
              for (idx_t r = 0; r < count; r++) {
                rev[r] = (sel_t)(count - 1 - r);
                ptrs2[r] = ptrs[count - 1 - r];
              }
              if (opt.lift_shape == 1) {
                one[0].Slice(SelectionVector(rev.data()));
              } else {
                size_t leaf_rows = 0;  // size the identity selection first: the slices keep pointers into it
                for (auto &kid : StructVector::GetEntries(one[0]))
                  if (kid->GetType().id() == LogicalTypeId::LIST && ListVector::GetEntry(*kid).GetType().id() == LogicalTypeId::LIST)
                    leaf_rows = std::max<size_t>(leaf_rows, ListVector::GetListSize(ListVector::GetEntry(*kid)));
                if (leaf_rows > ident.size()) {
                  const size_t old = ident.size();
                  ident.resize(leaf_rows);
                  for (size_t i = old; i < ident.size(); i++) ident[i] = (sel_t)i;
                }
                for (auto &kid : StructVector::GetEntries(one[0])) {
                  kid->Slice(SelectionVector(rev.data()));
                  if (kid->GetType().id() != LogicalTypeId::LIST) continue;
                  Vector &el = ListVector::GetEntry(*kid);
                  if (el.GetType().id() != LogicalTypeId::LIST) continue;
                  for (auto &leaf : StructVector::GetEntries(ListVector::GetEntry(el))) leaf->Slice(SelectionVector(ident.data()));
                }
              }
              Vector sv2(LogicalType::POINTER, (data_ptr_t)ptrs2.data());
              fun.update(one.data(), aggr, 1, sv2, count);
            } else if (opt.lift_shape == 3) {
              // the chunk's first lifted row as a CONSTANT vector `count` rows long
              one[0].SetVectorType(VectorType::CONSTANT_VECTOR);
              fun.update(one.data(), aggr, 1, state_vector, count);
            } else if (simple) {
              fun.simple_update(one.data(), aggr, 1, ptrs[0], count);
            } else {
              fun.update(one.data(), aggr, 1, state_vector, count);
            }
          } else if (simple) {
            fun.simple_update(chunk.data.data(), aggr, (idx_t)n_cols, ptrs[0], count);
          } else {
            fun.update(chunk.data.data(), aggr, (idx_t)n_cols, state_vector, count);
          }
        }
        // a worker that kept several state sets merges them into its first one (same thread, same arenas)
        for (int set = 1; set < K; set++) {
          std::vector<data_ptr_t> src_p, dst_p;
          for (int g = 0; g < n_groups; g++)
            if (loc.live[(size_t)set * n_groups + g]) {
              data_ptr_t d = loc.mem.get() + (size_t)g * ssz;
              if (!loc.live[g]) {
                fun.initialize(d);
                loc.live[g] = 1;
              }
              src_p.push_back(loc.mem.get() + ((size_t)set * n_groups + g) * ssz);
              dst_p.push_back(d);
            }
          if (src_p.empty()) continue;
          Vector sv(LogicalType::POINTER, (data_ptr_t)src_p.data()), dv(LogicalType::POINTER, (data_ptr_t)dst_p.data());
          fun.combine(sv, dv, aggr, src_p.size());
          if (fun.destructor) fun.destructor(sv, aggr, src_p.size());
        }
      } catch (std::exception &e) {
        std::lock_guard<std::mutex> g(err_mu);
        worker_error = e.what();
      }
    };
    if (T == 1) {
      worker(0);
    } else {
      std::vector<std::thread> pool;
      for (int t = 0; t < T; t++) pool.emplace_back(worker, t);
      for (auto &th : pool) th.join();
    }
    if (!worker_error.empty()) throw Exception(worker_error);

    // combine the thread-local states into thread 0's, group by group, then destroy the sources; finalize.
    // P > 1 threads do this for disjoint sets of groups at the same time, as DuckDB finalizes radix partitions.
    ThreadLocalStates &dst = tls[0];
    const int P = std::max(1, std::min(opt.parallel_finalize, n_groups));
    std::vector<std::vector<std::string>> rendered(P);  // per finalize thread: one STRUCT per live group, group order
    std::vector<std::vector<int>> groups_of(P);
    double secs = 0, finalize_secs = -1;
    double *seconds_box = &finalize_secs;
    auto finisher = [&](int p) {
      try {
        std::vector<int> mine;
        for (int g = p; g < n_groups; g += P) mine.push_back(g);
        for (int t = 1; t < T; t++) {
          std::vector<data_ptr_t> src_p, dst_p;
          for (int g : mine)
            if (tls[t].live[g]) {
              data_ptr_t d = dst.mem.get() + (size_t)g * ssz;
              if (!dst.live[g]) {
                fun.initialize(d);
                dst.live[g] = 1;
              }
              src_p.push_back(tls[t].mem.get() + (size_t)g * ssz);
              dst_p.push_back(d);
            }
          if (src_p.empty()) continue;
          Vector sv(LogicalType::POINTER, (data_ptr_t)src_p.data()), dv(LogicalType::POINTER, (data_ptr_t)dst_p.data());
          fun.combine(sv, dv, aggr, src_p.size());
          if (fun.destructor) fun.destructor(sv, aggr, src_p.size());
        }
        std::vector<data_ptr_t> fin;
        for (int g : mine)
          if (dst.live[g]) {
            fin.push_back(dst.mem.get() + (size_t)g * ssz);
            groups_of[p].push_back(g);
          }
        if (fin.empty()) return;
        Vector states(LogicalType::POINTER, (data_ptr_t)fin.data());
        Vector result(fun.return_type, std::max<idx_t>(fin.size(), 1));
        fun.finalize(states, aggr, result, fin.size(), 0);
        if (p == 0 && P == 1) *seconds_box = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        for (size_t i = 0; i < fin.size(); i++) {
          std::ostringstream one;
          RenderValue(result, i, one);
          rendered[p].push_back(one.str());
        }
        if (fun.destructor) fun.destructor(states, aggr, fin.size());
      } catch (std::exception &e) {
        std::lock_guard<std::mutex> g(err_mu);
        worker_error = e.what();
      }
    };
    if (P == 1) {
      finisher(0);
    } else {
      std::vector<std::thread> pool;
      for (int p = 0; p < P; p++) pool.emplace_back(finisher, p);
      for (auto &th : pool) th.join();
    }
    if (!worker_error.empty()) throw Exception(worker_error);
    secs = finalize_secs >= 0 ? finalize_secs : std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::ostringstream os;
    os << "[";
    {
      // results in group order
      std::vector<std::pair<int, const std::string *>> all;
      for (int p = 0; p < P; p++)
        for (size_t i = 0; i < groups_of[p].size(); i++) all.emplace_back(groups_of[p][i], &rendered[p][i]);
      std::sort(all.begin(), all.end(), [](const auto &x, const auto &y) { return x.first < y.first; });
      for (size_t i = 0; i < all.size(); i++) {
        if (i) os << ", ";
        os << *all[i].second;
      }
    }
    os << "]";
    if (seconds) *seconds = secs;
    *json_out = strdup(os.str().c_str());
    return 0;
  } catch (std::exception &e) {
    g_error = e.what();
    return -1;
  }
}

}  // extern "C"
