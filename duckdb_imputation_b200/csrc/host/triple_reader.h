// triple_reader.h -- one row of a ring STRUCT vector (any vector shape, read through TripleView) as a cfb_result:
// what the scalar functions over finished triples take in (multiply_triple, mul.cpp:24-60; the trainers' extract_data,
// ML/utils.cpp:4-150).
#pragma once
#include <cstring>
#include <vector>

#include "../../../include/cofactor_b200.h"
#include "triple_view.h"

namespace Triple {

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

}  // namespace Triple
