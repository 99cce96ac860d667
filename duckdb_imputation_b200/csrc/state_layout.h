// state_layout.h -- dense device layout of one aggregate context ("SumState" on the GPU).
//
// The reference keeps per state a float block and node-based maps (sum_state.h:14-28).  Here a
// state is two flat arrays per GROUP BY slot, indexed through dense key->slot tables
// (slot = key - lo[c]):
//   f64 [ lin n | quad nq | numcat n x total_dom ]      numcat index: i*total_dom + cat_off[c] + slot
//   u64 [ N | cat counts total_dom | pair counts ]      pair (k<l) index: pair_off[k*m+l] + slot_k*dom[l] + slot_l
// quad is the packed upper triangle (CFB_TRIPLE) or the diagonal (CFB_NB); numcat and the pair
// region are absent for CFB_NB.  Diagonal pairs (k,k) are not stored: they equal the cat counts.
#pragma once
#include <cstdint>

namespace cfb {

constexpr int kMaxCat = 32;

struct Layout {
  int kind, n, m, n_groups;
  int nq;                  // n(n+1)/2 or n
  int has_domain;          // 0 until a categorical domain has been fixed
  int pairs_hashed;        // 1: pair counts live in the PairHash (pair_hash.cuh), not in the u64 array
  int pad_;
  long long total_dom;     // sum of dom[c]
  long long F, U;          // per-group length of the f64 / u64 arrays
  long long numcat_base;   // n + nq
  long long pair_base;     // 1 + total_dom
  int lo[kMaxCat], dom[kMaxCat];
  long long cat_off[kMaxCat + 1];
  long long pair_off[kMaxCat * kMaxCat];  // [k*m + l], k < l, relative to pair_base
};

// Optional per-column slot translation for remaps between states whose columns are keyed
// differently (key dictionaries): dst slot = col[c] ? col[c][src slot] : src slot + (lo_src - lo_dst).
struct SlotTrans {
  const int *col[kMaxCat];
};

// Device column pointers of one scan (SoA input).
struct ScanCols {
  const float *num[32];
  const int32_t *cat[32];
  const int32_t *group;  // per-row GROUP BY slot or nullptr
};

}  // namespace cfb
