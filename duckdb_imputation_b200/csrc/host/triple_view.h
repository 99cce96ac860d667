// triple_view.h -- read a ring STRUCT vector (the output of to_cofactor / multiply_triple / a join over
// aggregate results) whatever its physical shape.
//
// The reference calls duckdb::RecursiveFlatten (utils.cpp:3-18) before touching the children with
// FlatVector::GetData (sum.cpp:72, mul.cpp:24-28): after a join or a filter DuckDB hands scalar functions
// and aggregates CONSTANT or DICTIONARY vectors, and a FLAT STRUCT vector may still have sliced
// (DICTIONARY) children.  Flattening copies every nested buffer; here every level is read through
// UnifiedVectorFormat instead (data[sel->get_index(i)]), which is the same contract without the copy:
//   STRUCT level   row r of the argument         -> sr = struct_sel(r) indexes every child
//   LIST level     list_entry_t of a child at sr -> (offset, length) into the list's child vector
//   leaf level     element e of a list child     -> data[leaf_sel(e)]
// Leaves that the C ABI wants as flat arrays (cfb_ctx_append_triples) are materialised only when
// they are not flat already (FlatLeaf).
#pragma once
#include <vector>

#include <duckdb.hpp>

namespace Triple {

// One vector at one nesting level, read through its unified format.
struct Level {
  duckdb::UnifiedVectorFormat fmt;
  Level() = default;
  Level(const Level &) = delete;  // fmt.sel may point into fmt itself
  Level &operator=(const Level &) = delete;
  void Bind(duckdb::Vector &v, idx_t count) { v.ToUnifiedFormat(count, fmt); }
  idx_t Index(idx_t i) const { return fmt.sel->get_index(i); }
  bool Flat() const { return fmt.sel->data() == nullptr; }
  template <class T>
  const T *Data() const {
    return duckdb::UnifiedVectorFormat::GetData<T>(fmt);
  }
  template <class T>
  T At(idx_t i) const {
    return Data<T>()[Index(i)];
  }
};

// LIST(T) child of the ring STRUCT: the list entries (indexed by the struct row) and the elements.
struct ListLevel {
  Level entries, elems;
  void Bind(duckdb::Vector &list, idx_t count) {
    entries.Bind(list, count);
    elems.Bind(duckdb::ListVector::GetEntry(list), duckdb::ListVector::GetListSize(list));
  }
  duckdb::list_entry_t Entry(idx_t sr) const { return entries.At<duckdb::list_entry_t>(sr); }
};

// LIST(LIST(STRUCT(k..., value))) child: outer entries per struct row, inner entries, the inner
// STRUCT's own selection, and its leaf vectors.
struct KeyValueLevel {
  Level outer, inner, rec;
  std::vector<Level> leaf;  // key[, key2], value
  idx_t n_elems = 0;
  void Bind(duckdb::Vector &list, idx_t count) {
    using namespace duckdb;
    outer.Bind(list, count);
    Vector &in = ListVector::GetEntry(list);
    inner.Bind(in, ListVector::GetListSize(list));
    Vector &st = ListVector::GetEntry(in);
    n_elems = ListVector::GetListSize(in);
    rec.Bind(st, n_elems);
    auto &kids = StructVector::GetEntries(st);
    leaf = std::vector<Level>(kids.size());
    for (size_t i = 0; i < kids.size(); i++) leaf[i].Bind(*kids[i], n_elems);
  }
  duckdb::list_entry_t Outer(idx_t sr) const { return outer.At<duckdb::list_entry_t>(sr); }
  duckdb::list_entry_t Inner(idx_t i) const { return inner.At<duckdb::list_entry_t>(i); }
  template <class T>
  T Leaf(size_t which, idx_t e) const {
    return leaf[which].At<T>(rec.Index(e));
  }
  // Leaf `which` as a flat array over [0, n_elems): the vector's own buffer when nothing is sliced,
  // else a gathered copy in `scratch`.
  template <class T>
  const T *FlatLeaf(size_t which, std::vector<T> &scratch) const {
    if (rec.Flat() && leaf[which].Flat()) return leaf[which].Data<T>();
    scratch.resize(n_elems);
    for (idx_t e = 0; e < n_elems; e++) scratch[e] = Leaf<T>(which, e);
    return scratch.data();
  }
};

// The whole ring STRUCT argument.
struct TripleView {
  bool nb;
  Level rows, N;
  ListLevel lin, quad;
  KeyValueLevel lin_cat, num_cat, cat_cat;

  TripleView(duckdb::Vector &v, idx_t count, bool nb_) : nb(nb_) {
    using namespace duckdb;
    if (v.GetType().id() != LogicalTypeId::STRUCT) throw InvalidInputException("expected a triple STRUCT");
    auto &kids = StructVector::GetEntries(v);
    if (kids.size() != (nb ? 4u : 6u)) throw InvalidInputException("triple STRUCT has the wrong number of fields");
    rows.Bind(v, count);
    // the children are indexed by the STRUCT's selection: bind them over every row it can name
    idx_t span = 0;
    for (idx_t r = 0; r < count; r++) span = std::max<idx_t>(span, rows.Index(r) + 1);
    N.Bind(*kids[0], span);
    lin.Bind(*kids[1], span);
    quad.Bind(*kids[2], span);
    lin_cat.Bind(*kids[3], span);
    if (!nb) {
      num_cat.Bind(*kids[4], span);
      cat_cat.Bind(*kids[5], span);
    }
  }
  idx_t Row(idx_t r) const { return rows.Index(r); }
};

}  // namespace Triple
