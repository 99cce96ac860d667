// pair_hash.cuh -- sparse (key1,key2) pair counts for large categorical domains.
//
// The dense pair tables of state_layout.h need D_k x D_l counters per column pair; for large
// domains (e.g. 10 columns of 10^5 keys: 4.5e11 counters) that is impossible, while the number
// of pairs that actually OCCUR is bounded by the number of rows -- the reference's
// std::map<std::pair<int,int>, float> (sum_state.h:27) is sparse for the same reason.  Above a
// size threshold the context therefore keeps pair counts in one open-addressing hash table in
// global memory: 64-bit key = (pair index : 9 | slot_k : 27 | slot_l : 27), linear probing,
// atomicCAS to claim a slot, 64-bit atomic add on the count; one table partition per GROUP BY slot.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace cfb {

constexpr unsigned long long kPairEmpty = ~0ull;
constexpr int kPairSlotBits = 27;

struct PairHash {
  unsigned long long *keys;       // [n_groups][capacity], kPairEmpty = free
  unsigned long long *counts;     // [n_groups][capacity]
  unsigned long long capacity;    // per group, power of two (0 = dense pair tables in use)
  unsigned long long *n_entries;  // device counter: occupied slots over all groups
};

__host__ __device__ __forceinline__ unsigned long long pair_key(int pair_index, long long sk, long long sl) {
  return ((unsigned long long)pair_index << (2 * kPairSlotBits)) | ((unsigned long long)sk << kPairSlotBits) |
         (unsigned long long)sl;
}
__host__ __device__ __forceinline__ unsigned long long pair_mix(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// counts[key] += inc in partition g; returns false when the partition is full.
__device__ __forceinline__ bool pair_hash_add(const PairHash &h, long long g, unsigned long long key, unsigned long long inc) {
  unsigned long long *keys = h.keys + g * h.capacity, *counts = h.counts + g * h.capacity;
  const unsigned long long mask = h.capacity - 1;
  unsigned long long i = pair_mix(key) & mask;
  for (unsigned long long probe = 0; probe < h.capacity; probe++, i = (i + 1) & mask) {
    unsigned long long cur = keys[i];
    if (cur == kPairEmpty) {
      cur = atomicCAS(keys + i, kPairEmpty, key);
      if (cur == kPairEmpty) {
        atomicAdd(h.n_entries, 1ull);
        cur = key;
      }
    }
    if (cur == key) {
      atomicAdd(counts + i, inc);
      return true;
    }
  }
  return false;
}

}  // namespace cfb
