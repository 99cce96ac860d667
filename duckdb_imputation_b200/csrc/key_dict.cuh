// key_dict.cuh -- per-column key dictionaries for categorical columns whose key RANGE is too wide
// for dense key->slot tables (slot = key - lo), e.g. ids scattered over the whole int32 range.
//
// The reference keeps std::map<int, ...> per column (sum_state.h:26), so any int32 is a valid key.
// Here such a column gets a device hash map key -> code, codes handed out in first-seen order; the
// scan kernels then run on the remapped column (a dense domain [0, n_codes)), and finalize maps
// codes back to keys and sorts them (std::map iteration order).  Two kernels per slice, no spinning:
//   dict_insert_kernel  claims a slot per new key (atomicCAS), the winner takes the next code and
//                       publishes it;
//   dict_remap_kernel   runs afterwards (all codes published) and rewrites keys to codes.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace cfb {

// An entry is (code << 32) | (uint32_t)key.  Real codes are < 2^31, so neither the free marker (code field
// 0xFFFFFFFF) nor the pending marker (0xFFFFFFFE) can be produced by a published entry, and
// (pending, key) differs from the free word for EVERY key, -1 included.
constexpr unsigned long long kDictEmpty = ~0ull;
constexpr unsigned int kDictPending = 0xFFFFFFFEu;

struct KeyDict {
  unsigned long long *table;  // [capacity]: (code << 32) | (uint32_t)key, kDictEmpty = free
  unsigned long long capacity;
  int *keys_of_code;          // [code_capacity]
  int *n_codes;               // device counter
};

__device__ __forceinline__ unsigned long long dict_hash(int key) {
  unsigned long long z = (unsigned long long)(unsigned int)key * 0x9E3779B97F4A7C15ull;
  return z ^ (z >> 29);
}

// Make sure `key` has a code.  Safe to call concurrently; never waits on another thread.
__device__ __forceinline__ void dict_insert(const KeyDict &d, int key) {
  const unsigned long long mask = d.capacity - 1;
  unsigned long long i = dict_hash(key) & mask;
  for (unsigned long long probe = 0; probe < d.capacity; probe++, i = (i + 1) & mask) {
    unsigned long long e = d.table[i];
    if (e == kDictEmpty) {
      const unsigned long long claim = ((unsigned long long)kDictPending << 32) | (unsigned int)key;
      e = atomicCAS(d.table + i, kDictEmpty, claim);
      if (e == kDictEmpty) {  // this thread owns the new key: assign and publish its code
        const int code = atomicAdd(d.n_codes, 1);
        d.keys_of_code[code] = key;
        atomicExch(d.table + i, ((unsigned long long)(unsigned int)code << 32) | (unsigned int)key);
        return;
      }
    }
    if ((int)(unsigned int)(e & 0xFFFFFFFFull) == key && (unsigned int)(e >> 32) != 0xFFFFFFFFu) return;  // present (its code may still be pending)
  }
}

// Code of a key that is known to be present with a published code; -1 if absent.
__device__ __forceinline__ int dict_lookup(const KeyDict &d, int key) {
  const unsigned long long mask = d.capacity - 1;
  unsigned long long i = dict_hash(key) & mask;
  for (unsigned long long probe = 0; probe < d.capacity; probe++, i = (i + 1) & mask) {
    const unsigned long long e = d.table[i];
    if (e == kDictEmpty) return -1;
    if ((int)(unsigned int)(e & 0xFFFFFFFFull) == key && (unsigned int)(e >> 32) != 0xFFFFFFFFu) return (int)(unsigned int)(e >> 32);
  }
  return -1;
}

__global__ void __launch_bounds__(256) dict_insert_kernel(KeyDict d, const int32_t *__restrict__ keys, unsigned long long n) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x)
    dict_insert(d, keys[i]);
}

__global__ void __launch_bounds__(256)
    dict_remap_kernel(KeyDict d, const int32_t *__restrict__ keys, int32_t *__restrict__ codes, unsigned long long n) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x)
    codes[i] = dict_lookup(d, keys[i]);
}

// Re-insert every entry of an old table into a larger one (codes are kept).
__global__ void __launch_bounds__(256) dict_rehash_kernel(KeyDict old_d, KeyDict new_d) {
  const unsigned long long mask = new_d.capacity - 1;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < old_d.capacity;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long e = old_d.table[i];
    if (e == kDictEmpty) continue;
    unsigned long long j = dict_hash((int)(unsigned int)(e & 0xFFFFFFFFull)) & mask;
    while (atomicCAS(new_d.table + j, kDictEmpty, e) != kDictEmpty) j = (j + 1) & mask;
  }
}

__global__ void dict_clear_kernel(unsigned long long *table, unsigned long long n) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x)
    table[i] = kDictEmpty;
}

// trans[s] = code in `dst` of the key that slot s of a source column stands for, inserting it if
// new.  Source keys: keys_src[s] (a source dictionary) or lo_src + s (a dense source column; then
// only slots with a non-zero count are translated, the rest get -1).  Two launches: insert, then
// (after the codes are published) lookup.
__global__ void __launch_bounds__(256)
    dict_translate_kernel(KeyDict dst, const int *__restrict__ keys_src, int lo_src, const unsigned long long *__restrict__ counts,
                          long long group_stride, int n_groups, long long n_slots, int *__restrict__ trans, int phase) {
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < n_slots; s += (long long)gridDim.x * blockDim.x) {
    bool seen = counts == nullptr;
    for (int g = 0; g < n_groups && !seen; g++) seen = counts[g * group_stride + s] != 0ull;
    if (!seen) {
      if (phase) trans[s] = -1;
      continue;
    }
    const int key = keys_src ? keys_src[s] : (int)((long long)lo_src + s);
    if (!phase)
      dict_insert(dst, key);
    else
      trans[s] = dict_lookup(dst, key);
  }
}

}  // namespace cfb
