// duckdb.hpp -- a minimal stand-in for the slice of the DuckDB 0.9.2 C++ API that the ring
// aggregates of eddbase/duckdb-imputation are written against.
//
// DuckDB is an un-vendored dependency of the reference (README.md:35-42) and is not installable
// in this environment, so this header is what lets (a) our callback glue (host/triple_glue.cpp)
// and (b) the reference's own sum_no_lift.cpp / sum_to_nb_agg.cpp / sum_state.cpp -- compiled
// unmodified from /root/reference into oracle/_ref -- be built and driven by the same
// hash-aggregate replay (host/replay_host.cpp).  With the real DuckDB headers on the include
// path instead of this directory the glue compiles against DuckDB proper.
//
// Only behaviour the callbacks rely on is modelled: flat / constant / dictionary vectors read
// through UnifiedVectorFormat, nested LIST / STRUCT result vectors with shared child buffers
// (Vector copies alias their buffers, as in DuckDB), ListVector::Reserve / SetListSize growth.
#pragma once
#include <cassert>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

typedef uint64_t idx_t;  // duckdb.h (C API) exposes idx_t globally; the reference relies on it

#ifndef DUCKDB_API
#define DUCKDB_API
#endif
#ifndef D_ASSERT
#define D_ASSERT(x) assert(x)
#endif
#define STANDARD_VECTOR_SIZE 2048
#ifndef PI
#define PI 3.141592653589793  // duckdb/common/constants: the reference's Box-Muller uses it (regression.cpp:502)
#endif

namespace duckdb {

using ::idx_t;
using std::string;
using data_t = uint8_t;
using data_ptr_t = data_t *;
using const_data_ptr_t = const data_t *;
using sel_t = uint32_t;
template <class T>
using vector = std::vector<T>;
template <class T>
using unique_ptr = std::unique_ptr<T>;
template <class T>
using shared_ptr = std::shared_ptr<T>;
template <class T, class... A>
unique_ptr<T> make_uniq(A &&...a) {
  return unique_ptr<T>(new T(std::forward<A>(a)...));
}
template <class T>
using child_list_t = std::vector<std::pair<std::string, T>>;

class Exception : public std::runtime_error {
 public:
  explicit Exception(const string &m) : std::runtime_error(m) {}
};
class InternalException : public Exception {
 public:
  explicit InternalException(const string &m) : Exception("INTERNAL Error: " + m) {}
};
class InvalidInputException : public Exception {
 public:
  explicit InvalidInputException(const string &m) : Exception("Invalid Input Error: " + m) {}
};

// --------------------------------------------------------------------------- types
enum class LogicalTypeId : uint8_t { INVALID, ANY, BOOLEAN, INTEGER, BIGINT, FLOAT, DOUBLE, VARCHAR, POINTER, STRUCT, LIST };
enum class PhysicalType : uint8_t { INVALID, BOOL, INT32, INT64, FLOAT, DOUBLE, VARCHAR, STRUCT, LIST };

class LogicalType {
 public:
  LogicalType() : id_(LogicalTypeId::INVALID) {}
  LogicalType(LogicalTypeId id) : id_(id) {}  // NOLINT: implicit like DuckDB's
  LogicalTypeId id() const { return id_; }
  PhysicalType InternalType() const {
    switch (id_) {
      case LogicalTypeId::BOOLEAN: return PhysicalType::BOOL;
      case LogicalTypeId::INTEGER: return PhysicalType::INT32;
      case LogicalTypeId::BIGINT: case LogicalTypeId::POINTER: return PhysicalType::INT64;
      case LogicalTypeId::FLOAT: return PhysicalType::FLOAT;
      case LogicalTypeId::DOUBLE: return PhysicalType::DOUBLE;
      case LogicalTypeId::STRUCT: return PhysicalType::STRUCT;
      case LogicalTypeId::LIST: return PhysicalType::LIST;
      default: return PhysicalType::INVALID;
    }
  }
  bool operator==(const LogicalType &o) const {
    if (id_ != o.id_ || children_.size() != o.children_.size()) return false;
    for (size_t i = 0; i < children_.size(); i++)
      if (children_[i].first != o.children_[i].first || !(children_[i].second == o.children_[i].second)) return false;
    return true;
  }
  bool operator!=(const LogicalType &o) const { return !(*this == o); }
  static LogicalType LIST(const LogicalType &child) {
    LogicalType t(LogicalTypeId::LIST);
    t.children_.emplace_back("", child);
    return t;
  }
  static LogicalType STRUCT(const child_list_t<LogicalType> &children) {
    LogicalType t(LogicalTypeId::STRUCT);
    t.children_ = children;
    return t;
  }
  const child_list_t<LogicalType> &children() const { return children_; }
  idx_t width() const {
    switch (id_) {
      case LogicalTypeId::BOOLEAN: return 1;
      case LogicalTypeId::INTEGER: case LogicalTypeId::FLOAT: return 4;
      case LogicalTypeId::BIGINT: case LogicalTypeId::DOUBLE: case LogicalTypeId::POINTER: return 8;
      case LogicalTypeId::LIST: return 16;  // list_entry_t
      default: return 0;                    // STRUCT has no data of its own
    }
  }
  static const LogicalType INTEGER, FLOAT, DOUBLE, BIGINT, BOOLEAN, ANY, POINTER, VARCHAR;

 private:
  LogicalTypeId id_;
  child_list_t<LogicalType> children_;
};
inline const LogicalType LogicalType::INTEGER = LogicalType(LogicalTypeId::INTEGER);
inline const LogicalType LogicalType::FLOAT = LogicalType(LogicalTypeId::FLOAT);
inline const LogicalType LogicalType::DOUBLE = LogicalType(LogicalTypeId::DOUBLE);
inline const LogicalType LogicalType::BIGINT = LogicalType(LogicalTypeId::BIGINT);
inline const LogicalType LogicalType::BOOLEAN = LogicalType(LogicalTypeId::BOOLEAN);
inline const LogicalType LogicalType::ANY = LogicalType(LogicalTypeId::ANY);
inline const LogicalType LogicalType::POINTER = LogicalType(LogicalTypeId::POINTER);
inline const LogicalType LogicalType::VARCHAR = LogicalType(LogicalTypeId::VARCHAR);

struct ListType {
  static const LogicalType &GetChildType(const LogicalType &t) { return t.children()[0].second; }
};
struct StructType {
  static const child_list_t<LogicalType> &GetChildTypes(const LogicalType &t) { return t.children(); }
};

struct list_entry_t {
  uint64_t offset;
  uint64_t length;
};

// ------------------------------------------------------------------------- vectors
enum class VectorType : uint8_t { FLAT_VECTOR, CONSTANT_VECTOR, DICTIONARY_VECTOR };

// DuckDB 0.9.2 spelling: the index array is private (`sel_vector`) and read through data(); a null
// array is the incremental selection (get_index(i) == i), which is what FLAT vectors report.
struct SelectionVector {
  SelectionVector() : sel_vector(nullptr) {}
  explicit SelectionVector(sel_t *s) : sel_vector(s) {}
  idx_t get_index(idx_t i) const { return sel_vector ? sel_vector[i] : i; }
  sel_t *data() { return sel_vector; }
  const sel_t *data() const { return sel_vector; }

 private:
  sel_t *sel_vector;
};

struct ValidityMask {
  bool RowIsValid(idx_t) const { return true; }  // the ring aggregates never read validity
  bool AllValid() const { return true; }
};

struct UnifiedVectorFormat {
  const SelectionVector *sel = nullptr;
  data_ptr_t data = nullptr;
  ValidityMask validity;
  SelectionVector owned_sel;
  template <class T>
  static const T *GetData(const UnifiedVectorFormat &f) {
    return reinterpret_cast<const T *>(f.data);
  }
};

// A scalar of one of the fixed-width types (what Vector::GetValue / SetValue exchange on this path), or a nested
// LIST / STRUCT value (what the MICE drivers' Value-level ring helpers take and return: duckdb::Value::LIST /
// ::STRUCT, ListValue::GetChildren, StructValue::GetChildren).
class Value {
 public:
  Value() : v_(0.0) {}
  Value(bool b) : v_(b ? 1.0 : 0.0) {}      // NOLINT: implicit like DuckDB's
  Value(int32_t i) : v_((double)i) {}       // NOLINT
  Value(int64_t i) : v_((double)i) {}       // NOLINT
  Value(float f) : v_((double)f) {}         // NOLINT
  Value(double d) : v_(d) {}                // NOLINT
  template <class T>
  T GetValue() const {
    if (nested_ != 0) throw Exception("Conversion Error: Unimplemented type for cast (nested value -> scalar)");
    return static_cast<T>(v_);
  }
  static Value LIST(const LogicalType &, vector<Value> values) {
    Value v;
    v.nested_ = 1;
    v.children_ = std::move(values);
    return v;
  }
  // DuckDB 0.9.2: the child type is taken from the first element, so an empty list is refused
  static Value LIST(vector<Value> values) {
    if (values.empty())
      throw InternalException("Value::LIST without providing a child-type requires a non-empty list of values. Use Value::LIST(child_type, list) instead.");
    return LIST(LogicalType(), std::move(values));
  }
  static Value STRUCT(child_list_t<Value> values) {
    Value v;
    v.nested_ = 2;
    for (auto &kv : values) {
      v.names_.push_back(kv.first);
      v.children_.push_back(std::move(kv.second));
    }
    return v;
  }
  bool IsList() const { return nested_ == 1; }
  bool IsStruct() const { return nested_ == 2; }
  const vector<Value> &NestedChildren() const { return children_; }
  const vector<string> &ChildNames() const { return names_; }

 private:
  double v_;
  int nested_ = 0;  // 0 scalar, 1 LIST, 2 STRUCT
  vector<Value> children_;
  vector<string> names_;
};
struct ListValue {
  static const vector<Value> &GetChildren(const Value &v) {
    if (!v.IsList()) throw InternalException("ListValue::GetChildren on a value that is not a LIST");
    return v.NestedChildren();
  }
};
struct StructValue {
  static const vector<Value> &GetChildren(const Value &v) {
    if (!v.IsStruct()) throw InternalException("StructValue::GetChildren on a value that is not a STRUCT");
    return v.NestedChildren();
  }
};

class Vector;

struct VectorBuffer {
  virtual ~VectorBuffer() = default;
  std::unique_ptr<data_t[]> bytes;
};
struct VectorListBuffer : VectorBuffer {
  unique_ptr<Vector> child;
  idx_t capacity = 0, size = 0;
};
struct VectorStructBuffer : VectorBuffer {
  vector<unique_ptr<Vector>> children;
};

class Vector {
 public:
  explicit Vector(LogicalType type, idx_t capacity = STANDARD_VECTOR_SIZE) : type_(std::move(type)) { Initialize(capacity); }
  // non-owning flat vector over caller memory (what a table scan hands to an aggregate)
  Vector(LogicalType type, data_ptr_t dataptr) : type_(std::move(type)), data_(dataptr) {}
  Vector(const Vector &) = default;  // aliases the buffers, like DuckDB's Vector(Vector &other)
  Vector &operator=(const Vector &) = default;

  const LogicalType &GetType() const { return type_; }
  VectorType GetVectorType() const { return vtype_; }
  void SetVectorType(VectorType t) { vtype_ = t; }
  data_ptr_t GetData() { return data_; }

  // DICTIONARY view of this (flat) vector through `sel` (caller keeps `sel` alive)
  void Slice(const SelectionVector &sel) {
    dict_sel_ = sel;
    vtype_ = VectorType::DICTIONARY_VECTOR;
  }

  void ToUnifiedFormat(idx_t count, UnifiedVectorFormat &f) const {
    (void)count;
    f.data = data_;
    switch (vtype_) {
      case VectorType::FLAT_VECTOR:
        f.owned_sel = SelectionVector(nullptr);
        break;
      case VectorType::DICTIONARY_VECTOR:
        f.owned_sel = dict_sel_;
        break;
      case VectorType::CONSTANT_VECTOR:
        f.owned_sel = SelectionVector(ZeroSel());
        break;
    }
    f.sel = &f.owned_sel;
  }

  // Materialise a CONSTANT / DICTIONARY vector of fixed-width values as FLAT (DuckDB's Vector::Flatten); nested
  // types keep their children (only the level's own data moves), which is all the callers here rely on.
  void Flatten(idx_t count) {
    if (vtype_ == VectorType::FLAT_VECTOR) return;
    const idx_t w = type_.width();
    if (w) {
      auto nb = std::make_shared<VectorBuffer>();
      nb->bytes.reset(new data_t[std::max<idx_t>(1, count * w)]());
      for (idx_t i = 0; i < count; i++) {
        const idx_t src = vtype_ == VectorType::CONSTANT_VECTOR ? 0 : dict_sel_.get_index(i);
        memcpy(nb->bytes.get() + i * w, data_ + src * w, w);
      }
      buffer_ = nb;
      data_ = nb->bytes.get();
    } else if (type_.id() == LogicalTypeId::STRUCT) {
      for (auto &c : static_cast<VectorStructBuffer &>(*aux_).children) {
        if (vtype_ == VectorType::CONSTANT_VECTOR) c->SetVectorType(VectorType::CONSTANT_VECTOR);
        else c->Slice(dict_sel_);
        c->Flatten(count);
      }
    }
    vtype_ = VectorType::FLAT_VECTOR;
    dict_sel_ = SelectionVector();
  }

  void Resize(idx_t cur, idx_t want) {
    if (type_.id() == LogicalTypeId::STRUCT) {
      for (auto &c : static_cast<VectorStructBuffer &>(*aux_).children) c->Resize(cur, want);
      return;
    }
    auto nb = std::make_shared<VectorBuffer>();
    const idx_t w = type_.width();
    nb->bytes.reset(new data_t[std::max<idx_t>(1, want * w)]());
    if (data_ && cur) memcpy(nb->bytes.get(), data_, cur * w);
    buffer_ = nb;
    data_ = nb->bytes.get();
  }

  shared_ptr<VectorBuffer> &auxiliary() { return aux_; }
  const shared_ptr<VectorBuffer> &auxiliary() const { return aux_; }

  // scalar access by row (fixed-width types only), through the vector's own selection
  Value GetValue(idx_t i) const {
    const idx_t r = vtype_ == VectorType::CONSTANT_VECTOR ? 0 : (vtype_ == VectorType::DICTIONARY_VECTOR ? dict_sel_.get_index(i) : i);
    switch (type_.id()) {
      case LogicalTypeId::BOOLEAN: return Value(reinterpret_cast<const uint8_t *>(data_)[r] != 0);
      case LogicalTypeId::INTEGER: return Value(reinterpret_cast<const int32_t *>(data_)[r]);
      case LogicalTypeId::BIGINT: return Value(reinterpret_cast<const int64_t *>(data_)[r]);
      case LogicalTypeId::FLOAT: return Value(reinterpret_cast<const float *>(data_)[r]);
      case LogicalTypeId::DOUBLE: return Value(reinterpret_cast<const double *>(data_)[r]);
      default: throw InternalException("shim: GetValue on a nested vector");
    }
  }
  void SetValue(idx_t i, const Value &v) {
    switch (type_.id()) {
      case LogicalTypeId::BOOLEAN: reinterpret_cast<uint8_t *>(data_)[i] = v.GetValue<bool>(); break;
      case LogicalTypeId::INTEGER: reinterpret_cast<int32_t *>(data_)[i] = v.GetValue<int32_t>(); break;
      case LogicalTypeId::BIGINT: reinterpret_cast<int64_t *>(data_)[i] = v.GetValue<int64_t>(); break;
      case LogicalTypeId::FLOAT: reinterpret_cast<float *>(data_)[i] = v.GetValue<float>(); break;
      case LogicalTypeId::DOUBLE: reinterpret_cast<double *>(data_)[i] = v.GetValue<double>(); break;
      default: throw InternalException("shim: SetValue on a nested vector");
    }
  }

 private:
  static sel_t *ZeroSel() {
    static sel_t zeros[STANDARD_VECTOR_SIZE] = {0};
    return zeros;
  }
  void Initialize(idx_t capacity) {
    const idx_t w = type_.width();
    if (w) {
      buffer_ = std::make_shared<VectorBuffer>();
      buffer_->bytes.reset(new data_t[std::max<idx_t>(1, capacity * w)]());
      data_ = buffer_->bytes.get();
    }
    if (type_.id() == LogicalTypeId::LIST) {
      auto lb = std::make_shared<VectorListBuffer>();
      lb->child = make_uniq<Vector>(ListType::GetChildType(type_), capacity);
      lb->capacity = capacity;
      aux_ = lb;
    } else if (type_.id() == LogicalTypeId::STRUCT) {
      auto sb = std::make_shared<VectorStructBuffer>();
      for (auto &c : type_.children()) sb->children.push_back(make_uniq<Vector>(c.second, capacity));
      aux_ = sb;
    }
  }
  LogicalType type_;
  VectorType vtype_ = VectorType::FLAT_VECTOR;
  data_ptr_t data_ = nullptr;
  shared_ptr<VectorBuffer> buffer_, aux_;
  SelectionVector dict_sel_;
};

struct FlatVector {
  template <class T>
  static T *GetData(Vector &v) {
    return reinterpret_cast<T *>(v.GetData());
  }
  static data_ptr_t GetData(Vector &v) { return v.GetData(); }
};
struct ConstantVector {
  template <class T>
  static T *GetData(Vector &v) {
    return reinterpret_cast<T *>(v.GetData());
  }
};

struct ListVector {
  static VectorListBuffer &Buf(Vector &v) { return static_cast<VectorListBuffer &>(*v.auxiliary()); }
  static Vector &GetEntry(Vector &v) { return *Buf(v).child; }
  static list_entry_t *GetData(Vector &v) { return reinterpret_cast<list_entry_t *>(v.GetData()); }
  static idx_t GetListSize(Vector &v) { return Buf(v).size; }
  static void SetListSize(Vector &v, idx_t n) { Buf(v).size = n; }
  static void Reserve(Vector &v, idx_t n) {
    auto &b = Buf(v);
    if (n > b.capacity) {
      idx_t cap = b.capacity ? b.capacity : 1;
      while (cap < n) cap *= 2;
      b.child->Resize(b.capacity, cap);
      b.capacity = cap;
    }
  }
};
struct StructVector {
  static vector<unique_ptr<Vector>> &GetEntries(Vector &v) {
    return static_cast<VectorStructBuffer &>(*v.auxiliary()).children;
  }
  static const vector<unique_ptr<Vector>> &GetEntries(const Vector &v) {
    return static_cast<const VectorStructBuffer &>(*v.auxiliary()).children;
  }
};

// --------------------------------------------------------------- function plumbing
class ClientContext {};
class DatabaseInstance;
class Expression {
 public:
  LogicalType return_type;
};
class BaseStatistics {  // only produced by the statistics callbacks, which nothing here calls
 public:
  unique_ptr<BaseStatistics> ToUnique() const { return make_uniq<BaseStatistics>(*this); }
};
struct NumericStats {
  static BaseStatistics CreateUnknown(const LogicalType &) { return BaseStatistics(); }
};
struct FunctionData {
  virtual ~FunctionData() = default;
  template <class T>
  T &Cast() {
    return reinterpret_cast<T &>(*this);
  }
};
struct FunctionLocalState {
  virtual ~FunctionLocalState() = default;
  template <class T>
  T &Cast() {
    return reinterpret_cast<T &>(*this);
  }
};
class BoundFunctionExpression : public Expression {};
struct FunctionStatisticsInput {
  BoundFunctionExpression &expr;
  vector<BaseStatistics> &child_stats;
};
struct VariableReturnBindData : FunctionData {
  explicit VariableReturnBindData(LogicalType t) : stype(std::move(t)) {}
  LogicalType stype;
};
struct AggregateInputData {
  AggregateInputData() = default;
  explicit AggregateInputData(FunctionData *bd) : bind_data(bd) {}
  FunctionData *bind_data = nullptr;
};
enum class FunctionNullHandling : uint8_t { DEFAULT_NULL_HANDLING, SPECIAL_HANDLING };

class AggregateFunction;
typedef idx_t (*aggregate_size_t)();
typedef void (*aggregate_initialize_t)(data_ptr_t state);
typedef void (*aggregate_update_t)(Vector inputs[], AggregateInputData &, idx_t input_count, Vector &state, idx_t count);
typedef void (*aggregate_combine_t)(Vector &state, Vector &combined, AggregateInputData &, idx_t count);
typedef void (*aggregate_finalize_t)(Vector &state, AggregateInputData &, Vector &result, idx_t count, idx_t offset);
typedef void (*aggregate_simple_update_t)(Vector inputs[], AggregateInputData &, idx_t input_count, data_ptr_t state, idx_t count);
typedef unique_ptr<FunctionData> (*bind_aggregate_function_t)(ClientContext &, AggregateFunction &, vector<unique_ptr<Expression>> &);
typedef void (*aggregate_destructor_t)(Vector &state, AggregateInputData &, idx_t count);
typedef void *aggregate_statistics_t;
typedef void *aggregate_window_t;

class AggregateFunction {
 public:
  AggregateFunction(string name_p, vector<LogicalType> arguments_p, LogicalType return_type_p, aggregate_size_t state_size_p,
                    aggregate_initialize_t initialize_p, aggregate_update_t update_p, aggregate_combine_t combine_p,
                    aggregate_finalize_t finalize_p, aggregate_simple_update_t simple_update_p = nullptr,
                    bind_aggregate_function_t bind_p = nullptr, aggregate_destructor_t destructor_p = nullptr,
                    aggregate_statistics_t statistics_p = nullptr, aggregate_window_t window_p = nullptr)
      : name(std::move(name_p)), arguments(std::move(arguments_p)), return_type(std::move(return_type_p)),
        state_size(state_size_p), initialize(initialize_p), update(update_p), combine(combine_p), finalize(finalize_p),
        simple_update(simple_update_p), bind(bind_p), destructor(destructor_p) {
    (void)statistics_p;
    (void)window_p;
  }
  string name;
  vector<LogicalType> arguments;
  LogicalType return_type;
  LogicalType varargs;
  FunctionNullHandling null_handling = FunctionNullHandling::DEFAULT_NULL_HANDLING;
  aggregate_size_t state_size;
  aggregate_initialize_t initialize;
  aggregate_update_t update;
  aggregate_combine_t combine;
  aggregate_finalize_t finalize;
  aggregate_simple_update_t simple_update;
  bind_aggregate_function_t bind;
  aggregate_destructor_t destructor;

  template <class STATE>
  static idx_t StateSize() {
    return sizeof(STATE);
  }
  template <class STATE, class OP>
  static void StateInitialize(data_ptr_t state) {
    OP::Initialize(*reinterpret_cast<STATE *>(state));
  }
  template <class STATE, class OP>
  static void StateDestroy(Vector &states, AggregateInputData &aggr_input_data, idx_t count) {
    auto sdata = FlatVector::GetData<STATE *>(states);
    for (idx_t i = 0; i < count; i++) OP::template Destroy<STATE>(*sdata[i], aggr_input_data);
  }
};

// ---- scalar functions (to_cofactor / to_nb_agg)
class DataChunk {
 public:
  vector<Vector> data;
  idx_t size() const { return count; }
  void SetCardinality(idx_t c) { count = c; }
  idx_t ColumnCount() const { return data.size(); }

 private:
  idx_t count = 0;
};
struct ExpressionState {
  unique_ptr<FunctionLocalState> local;  // what init_local_state returned for this executor
};
struct ExecuteFunctionState {
  static FunctionLocalState *GetFunctionState(ExpressionState &state) { return state.local.get(); }
};
class ScalarFunction;
typedef void (*scalar_function_t)(DataChunk &args, ExpressionState &state, Vector &result);
typedef unique_ptr<FunctionData> (*bind_scalar_function_t)(ClientContext &, ScalarFunction &, vector<unique_ptr<Expression>> &);
typedef unique_ptr<BaseStatistics> (*function_statistics_t)(ClientContext &, FunctionStatisticsInput &);
typedef unique_ptr<FunctionLocalState> (*init_local_state_t)(ExpressionState &, const BoundFunctionExpression &, FunctionData *);
class ScalarFunction {
 public:
  ScalarFunction(string name_p, vector<LogicalType> arguments_p, LogicalType return_type_p, scalar_function_t function_p,
                 bind_scalar_function_t bind_p = nullptr, void *dependency = nullptr, function_statistics_t statistics_p = nullptr,
                 init_local_state_t init_local_state_p = nullptr)
      : name(std::move(name_p)), arguments(std::move(arguments_p)), return_type(std::move(return_type_p)),
        function(function_p), bind(bind_p), statistics(statistics_p), init_local_state(init_local_state_p) {
    (void)dependency;
  }

  string name;
  vector<LogicalType> arguments;
  LogicalType return_type;
  LogicalType varargs;
  FunctionNullHandling null_handling = FunctionNullHandling::DEFAULT_NULL_HANDLING;
  scalar_function_t function;
  bind_scalar_function_t bind;
  function_statistics_t statistics;
  init_local_state_t init_local_state;
};

// The catalog an extension registers into (ExtensionUtil::RegisterFunction).
class DatabaseInstance {
 public:
  std::map<string, AggregateFunction> aggregates;
  std::map<string, ScalarFunction> scalars;
};

// What a loadable extension's entry points touch: duckdb_<name>_init(DatabaseInstance&) wraps the instance in a
// DuckDB object and calls LoadExtension<T>(), duckdb_<name>_version() returns DuckDB::LibraryVersion()
// (duckdb_imputation_extension.cpp:269-279).
#ifndef DUCKDB_EXTENSION_API
#define DUCKDB_EXTENSION_API __attribute__((visibility("default")))
#endif
class DuckDB;
class Extension {
 public:
  virtual ~Extension() = default;
  virtual void Load(DuckDB &db) = 0;
  virtual std::string Name() = 0;
};
class DuckDB {
 public:
  explicit DuckDB(DatabaseInstance &i) : instance(&i, [](DatabaseInstance *) {}) {}
  template <class T>
  void LoadExtension() {
    T extension;
    extension.Load(*this);
  }
  static const char *LibraryVersion() { return "v0.9.2"; }  // the DuckDB release the reference pins (README.md:35-42)
  shared_ptr<DatabaseInstance> instance;
};
struct ExtensionUtil {
  static void RegisterFunction(DatabaseInstance &db, AggregateFunction f) {
    const string name = f.name;
    db.aggregates.erase(name);
    db.aggregates.emplace(name, std::move(f));
  }
  static void RegisterFunction(DatabaseInstance &db, ScalarFunction f) {
    const string name = f.name;
    db.scalars.erase(name);
    db.scalars.emplace(name, std::move(f));
  }
};

}  // namespace duckdb
