// cofactor_b200.cu -- the C ABI of include/cofactor_b200.h over the sm_100a kernels.
//
// One cfb_ctx is one aggregate state ("SumState" of sum_state.h:14-28 shrunk to a handle):
// a dense device-resident state (state_layout.h), a stream, the scratch of the Gram kernel
// and a double-buffered pinned staging ring for host (DuckDB vector) input.
// There is no CPU fallback anywhere in this file: without a CUDA device every compute entry
// point returns CFB_ERR_NO_DEVICE.
#include "../../include/cofactor_b200.h"

#include <algorithm>
#include <array>
#include <atomic>
#include <climits>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <mutex>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <emmintrin.h>
#include <nccl.h>
#include <nvtx3/nvToolsExt.h>  // header-only; a no-op unless a profiler injects itself

#include "gram_launch.h"
#include "group_kernel.cuh"
#include "key_count_kernel.cuh"
#include "key_dict.cuh"
#include "pair_hash.cuh"
#include "predict_kernel.cuh"
#include "slab_kernels.cuh"
#include "slot_gram_kernel.cuh"
#include "slot_block_kernel.cuh"
#include "bucket_kernels.cuh"
#include "bucket_launch.h"
#include "chain_kernels.cuh"
#include "chain_launch.h"
#include "role_kernels.cuh"
#include "role_launch.h"
#include "slab_launch.h"
#include "scatter_kernels.cuh"
#include "state_layout.h"

using cfb::Layout;

namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_timing{0};

// One NVTX range per C-ABI call (tracing): nsys / ncu --nvtx show the callbacks' appends, combines and
// finalizes over the kernels they launch.
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange &) = delete;
  NvtxRange &operator=(const NvtxRange &) = delete;
};

int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(call)                                                                                        \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess)                                                                              \
      return fail(e_ == cudaErrorMemoryAllocation ? CFB_ERR_OOM : CFB_ERR_CUDA, "%s failed: %s (%s:%d)", \
                  #call, cudaGetErrorString(e_), __FILE__, __LINE__);                                   \
  } while (0)

int device_count_quiet() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

struct DeviceInfo {
  int sms = 0;
  int smem_optin = 0;
  int smem_sm = 0;  // shared memory of one SM (all resident CTAs together)
};
DeviceInfo g_dev[64];
std::once_flag g_dev_once[64];

const DeviceInfo &dev_info(int d) {
  std::call_once(g_dev_once[d], [d] {
    cudaDeviceGetAttribute(&g_dev[d].sms, cudaDevAttrMultiProcessorCount, d);
    cudaDeviceGetAttribute(&g_dev[d].smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, d);
    cudaDeviceGetAttribute(&g_dev[d].smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, d);
  });
  return g_dev[d];
}

struct Stage {
  char *h = nullptr;  // pinned: [col][tile_rows] x 4 B, numeric cols, cat cols, group slot
  char *d = nullptr;
  size_t bytes = 0;
  cudaEvent_t done = nullptr;
  bool in_flight = false;
};

// Staging buffers (pinned host + device twin) are recycled across contexts: pinning memory costs
// milliseconds, and a DuckDB query creates one state per worker thread (and per group).  Three
// size classes: a lone aggregate gets 16 MB tiles (PCIe-efficient), a GROUP BY with hundreds of
// live states gets 1 MB tiles so that pinned memory stays bounded.
constexpr size_t kStageBytesClass[3] = {16u << 20, 4u << 20, 1u << 20};
struct StagePool {
  std::mutex mu;
  std::vector<Stage> free_list[64][3];
  int live[64] = {0};  // stages handed out per device
  static int class_of(size_t bytes) { return bytes == kStageBytesClass[0] ? 0 : (bytes == kStageBytesClass[1] ? 1 : 2); }
  int acquire(int device, Stage *out) {
    int cls;
    {
      std::lock_guard<std::mutex> g(mu);
      const int n = live[device & 63];
      cls = n < 32 ? 0 : (n < 256 ? 1 : 2);
      live[device & 63]++;
      auto &fl = free_list[device & 63][cls];
      if (!fl.empty()) {
        *out = fl.back();
        fl.pop_back();
        return CFB_OK;
      }
    }
    Stage s;
    s.bytes = kStageBytesClass[cls];
    cudaError_t e = cudaHostAlloc((void **)&s.h, s.bytes, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaMalloc((void **)&s.d, s.bytes);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming);
    if (e != cudaSuccess) {
      if (s.h) cudaFreeHost(s.h);
      if (s.d) cudaFree(s.d);
      std::lock_guard<std::mutex> g(mu);
      live[device & 63]--;
      return fail(e == cudaErrorMemoryAllocation ? CFB_ERR_OOM : CFB_ERR_CUDA, "staging allocation: %s", cudaGetErrorString(e));
    }
    *out = s;
    return CFB_OK;
  }
  void release(int device, Stage &s) {
    if (!s.h) return;
    s.in_flight = false;
    std::lock_guard<std::mutex> g(mu);
    live[device & 63]--;
    free_list[device & 63][class_of(s.bytes)].push_back(s);
    s = Stage{};
  }
};
StagePool g_stage_pool;

}  // namespace

struct cfb_ctx {
  int device = 0, kind = 0, n = 0, m = 0, G = 1;
  bool user_domain = false;  // set through cfb_ctx_set_cat_domain: out-of-range keys are errors
  bool touched = false;      // something was aggregated since creation / recycling
  Layout lay{};
  Layout *d_lay = nullptr;
  double *d_f64 = nullptr;
  unsigned long long *d_u64 = nullptr;
  int *d_err = nullptr;
  cfb::PairHash hash{};               // sparse pair counts (lay.pairs_hashed)
  struct ColDict {                    // key dictionary of a wide-range categorical column (key_dict.cuh)
    bool on = false;
    cfb::KeyDict d{};
    unsigned long long code_cap = 0;
    int n_codes = 0;                  // host copy, exact after the last dictionary pass
  } dict[cfb::kMaxCat];
  int32_t *d_remap = nullptr;         // [dict columns][remap_rows] keys rewritten to codes (device scans)
  size_t remap_rows = 0, remap_cap = 0;
  unsigned long long hash_upper = 0;  // host-side upper bound on occupied hash slots
  cudaStream_t stream = nullptr;
  cudaStream_t user_stream = nullptr;  // last caller-provided stream of cfb_triple_device
  // per-CTA fp32 slabs of slab_scan_kernel (all zero between launches)
  float *d_slab = nullptr;
  long long slab_floats = 0;  // per CTA
  int slab_grid = 0;
  // chain_sum_kernel: per-CTA count slabs (all zero between launches) and the packed one-byte slots it hands to
  // pair_packed_kernel
  float *d_chain_slab = nullptr;  // its own fp32 slab: sharing d_slab with the slab / role kernels re-allocated it (with a
  long long chain_slab_floats = 0;  // stream sync) every time a context alternated between large and small scans
  int chain_slab_grid = 0;
  unsigned *d_cnt_slab = nullptr;
  long long cnt_slab_words = 0;  // per CTA
  int cnt_slab_grid = 0;
  unsigned char *d_packed = nullptr;
  size_t packed_cap = 0;
  // role plan of role_scan_kernel (shared-memory pair tables), rebuilt when the domains change
  cfb::RolePlan *role_plan = nullptr;  // host copy; travels in the kernel parameters
  int role_state = 0;  // 0: no plan for the current domains yet, 1: plan valid, -1: shape does not fit
  int role_dom[cfb::kMaxCat] = {0};
  int role_roles = 0, role_bits = 0;
  size_t role_smem = 0;
  // Gram scratch
  double *d_partials = nullptr;
  unsigned int *d_ticket = nullptr;
  int gram_grid = 0;
  // timing of the most recent device scan
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
  // staging ring for host input
  Stage st[2];
  size_t tile_rows = 0, fill = 0;
  int cur = 0;
  bool uses_group = false;
  int st_lo[cfb::kMaxCat], st_hi[cfb::kMaxCat];  // min/max of the keys staged in the open tile
  // A state without a dense partial (hashed pair counts, key dictionaries, a domain every rank discovered for
  // itself) is all-reduced as results: per group the global result lives here, finalize hands out copies, and
  // the context accepts no further input (its device state is still the local one).
  std::vector<cfb_result> reduced;
};

namespace {

// Contexts are recycled: a DuckDB query creates one state per worker thread (and per group) and
// destroys them all at the end; cudaMalloc / cudaFree / stream creation cost milliseconds and
// serialise inside the driver, so a destroyed context keeps its stream, events and device
// buffers, is zeroed, and waits here for the next cfb_ctx_create of the same shape.
struct CtxPool {
  std::mutex mu;
  std::vector<cfb_ctx *> idle;
  static constexpr size_t kMaxIdle = 128;
  static constexpr long long kMaxStateBytes = 8ll << 20;
  cfb_ctx *take(int device, int kind, int n, int m, int G) {
    std::lock_guard<std::mutex> g(mu);
    for (size_t i = 0; i < idle.size(); i++) {
      cfb_ctx *c = idle[i];
      if (c->device == device && c->kind == kind && c->n == n && c->m == m && c->G == G) {
        idle[i] = idle.back();
        idle.pop_back();
        return c;
      }
    }
    return nullptr;
  }
  bool give(cfb_ctx *c) {
    std::lock_guard<std::mutex> g(mu);
    if (idle.size() >= kMaxIdle) return false;
    idle.push_back(c);
    return true;
  }
};
CtxPool g_ctx_pool;

int check_open(const cfb_ctx *c) {
  if (c && !c->reduced.empty())
    return fail(CFB_ERR_STATE, "this context holds an all-reduced result (cfb_ctx_allreduce of a sparse state): finalize it, it takes no more input");
  return CFB_OK;
}
// deep copy of a result (every array malloc'd, as cfb_result_free expects)
int copy_result(const cfb_result *a, cfb_result *out) {
  *out = *a;
  auto dup = [](const void *p, size_t bytes) {
    void *q = malloc(std::max<size_t>(8, bytes));
    if (q && p && bytes) memcpy(q, p, bytes);
    return q;
  };
  const size_t tk = (size_t)a->total_keys, np = (size_t)(a->n_pair_lists ? a->pair_offsets[a->n_pair_lists] : 0);
  out->lin = (double *)dup(a->lin, (size_t)a->n_num * 8);
  out->quad = (double *)dup(a->quad, (size_t)a->n_quad * 8);
  out->cat_offsets = (int64_t *)dup(a->cat_offsets, ((size_t)a->n_cat + 1) * 8);
  out->cat_keys = (int32_t *)dup(a->cat_keys, tk * 4);
  out->cat_counts = (int64_t *)dup(a->cat_counts, tk * 8);
  out->numcat_sums = a->numcat_sums ? (double *)dup(a->numcat_sums, (size_t)a->n_num * tk * 8) : nullptr;
  out->pair_offsets = (int64_t *)dup(a->pair_offsets, ((size_t)a->n_pair_lists + 1) * 8);
  out->pair_key1 = (int32_t *)dup(a->pair_key1, np * 4);
  out->pair_key2 = (int32_t *)dup(a->pair_key2, np * 4);
  out->pair_counts = (int64_t *)dup(a->pair_counts, np * 8);
  return CFB_OK;
}
void drop_reduced(cfb_ctx *c) {
  for (auto &r : c->reduced) cfb_result_free(&r);
  c->reduced.clear();
}

// ------------------------------------------------------------------------- layout
// Dense pair tables are used while they stay below this many bytes (all groups); above it the
// pair counts go to the hash table of pair_hash.cuh.  CFB_DENSE_PAIR_BYTES overrides (tests).
long long dense_pair_limit() {
  static const long long v = [] {
    const char *e = getenv("CFB_DENSE_PAIR_BYTES");
    return e ? atoll(e) : (2ll << 30);
  }();
  return v;
}

void build_layout(Layout &L, int kind, int n, int m, int G, const int *lo, const int *hi) {
  memset(&L, 0, sizeof(L));
  L.kind = kind;
  L.n = n;
  L.m = m;
  L.n_groups = G;
  L.nq = kind == CFB_NB ? n : n * (n + 1) / 2;
  L.has_domain = (lo != nullptr) || m == 0;
  long long off = 0;
  for (int c = 0; c < m; c++) {
    L.lo[c] = lo ? lo[c] : 0;
    L.dom[c] = lo ? (int)((long long)hi[c] - lo[c] + 1) : 0;
    L.cat_off[c] = off;
    off += L.dom[c];
  }
  L.cat_off[m] = off;
  L.total_dom = off;
  L.numcat_base = n + L.nq;
  L.pair_base = 1 + off;
  long long po = 0;
  if (kind == CFB_TRIPLE)
    for (int k = 0; k < m; k++)
      for (int l = k + 1; l < m; l++) {
        L.pair_off[k * m + l] = po;
        po += (long long)L.dom[k] * L.dom[l];
      }
  if (po * 8 * G > dense_pair_limit()) {
    L.pairs_hashed = 1;  // sparse pair counts (pair_hash.cuh)
    po = 0;
  }
  L.F = L.numcat_base + (kind == CFB_TRIPLE ? (long long)n * off : 0);
  L.U = L.pair_base + po;
}

constexpr long long kMaxStateBytes = 16ll << 30;

int alloc_state(const Layout &L, double **f, unsigned long long **u, cudaStream_t s) {
  const long long bf = L.F * L.n_groups * 8, bu = L.U * L.n_groups * 8;
  if (bf + bu > kMaxStateBytes)
    return fail(CFB_ERR_DOMAIN, "dense categorical state would need %lld bytes (limit %lld): domain too large",
                bf + bu, kMaxStateBytes);
  CU(cudaMalloc(f, std::max<long long>(bf, 8)));
  CU(cudaMalloc(u, std::max<long long>(bu, 8)));
  CU(cudaMemsetAsync(*f, 0, std::max<long long>(bf, 8), s));
  CU(cudaMemsetAsync(*u, 0, std::max<long long>(bu, 8), s));
  return CFB_OK;
}

// ---- sparse pair counts ------------------------------------------------------------
constexpr unsigned long long kMinHashCapacity = 1ull << 16, kMaxHashCapacity = 1ull << 30;

// Large hash tables (hundreds of MB to GB) wait here for the next context / the next doubling instead of going back
// to the driver: cudaMalloc + cudaFree of that size inside a scan cost tens to hundreds of milliseconds.  A pooled
// pair is handed out only after the stream that last used it was drained (callers synchronise before hash_free).
struct HashPool {
  std::mutex mu;
  struct Buf {
    int device;
    size_t bytes;
    unsigned long long *keys, *counts;
  };
  std::vector<Buf> idle;
  static constexpr size_t kMaxIdle = 3, kMinBytes = 32u << 20;
  bool take(int device, size_t bytes, unsigned long long **keys, unsigned long long **counts) {
    std::lock_guard<std::mutex> g(mu);
    for (size_t i = 0; i < idle.size(); i++)
      if (idle[i].device == device && idle[i].bytes == bytes) {
        *keys = idle[i].keys;
        *counts = idle[i].counts;
        idle[i] = idle.back();
        idle.pop_back();
        return true;
      }
    return false;
  }
  bool give(int device, size_t bytes, unsigned long long *keys, unsigned long long *counts) {
    if (bytes < kMinBytes) return false;
    std::lock_guard<std::mutex> g(mu);
    if (idle.size() >= kMaxIdle) {  // drop the smallest
      size_t smallest = 0;
      for (size_t i = 1; i < idle.size(); i++)
        if (idle[i].bytes < idle[smallest].bytes) smallest = i;
      if (idle[smallest].bytes >= bytes) return false;
      cudaFree(idle[smallest].keys);
      cudaFree(idle[smallest].counts);
      idle[smallest] = idle.back();
      idle.pop_back();
    }
    idle.push_back({device, bytes, keys, counts});
    return true;
  }
};
HashPool g_hash_pool;

void hash_free(cfb::PairHash &h, int G = 0) {
  int device = 0;
  cudaGetDevice(&device);
  const size_t bytes = (size_t)h.capacity * (size_t)std::max(G, 0) * 8;
  if (!(G > 0 && h.keys && h.counts && g_hash_pool.give(device, bytes, h.keys, h.counts))) {
    cudaFree(h.keys);
    cudaFree(h.counts);
  }
  cudaFree(h.n_entries);
  h = cfb::PairHash{};
}

int hash_alloc(cfb::PairHash *h, unsigned long long capacity, int G, cudaStream_t s) {
  *h = cfb::PairHash{};
  if (capacity > kMaxHashCapacity) return fail(CFB_ERR_DOMAIN, "pair hash table would need %llu slots per group", capacity);
  h->capacity = capacity;
  const size_t bytes = (size_t)capacity * G * 8;
  int device = 0;
  cudaGetDevice(&device);
  cudaError_t e = cudaSuccess;
  if (!g_hash_pool.take(device, bytes, &h->keys, &h->counts)) {
    e = cudaMalloc(&h->keys, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&h->counts, bytes);
  }
  if (e == cudaSuccess) e = cudaMalloc(&h->n_entries, 8);
  if (e != cudaSuccess) {
    hash_free(*h);
    return fail(e == cudaErrorMemoryAllocation ? CFB_ERR_OOM : CFB_ERR_CUDA, "pair hash allocation: %s", cudaGetErrorString(e));
  }
  cfb::pair_hash_clear_kernel<<<148 * 4, 256, 0, s>>>(*h, capacity * G);
  g_launches++;
  CU(cudaGetLastError());
  return CFB_OK;
}

unsigned long long pow2_at_least(unsigned long long v) {
  unsigned long long p = kMinHashCapacity;
  while (p < v) p <<= 1;
  return p;
}

// Make room for `add` more distinct pairs (per group, worst case) before a scan.
int hash_reserve(cfb_ctx *c, unsigned long long add) {
  if (!c->lay.pairs_hashed) return CFB_OK;
  if ((c->hash_upper + add) * 2 <= c->hash.capacity) {
    c->hash_upper += add;
    return CFB_OK;
  }
  unsigned long long exact = 0;
  CU(cudaMemcpyAsync(&exact, c->hash.n_entries, 8, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  c->hash_upper = exact;
  if ((exact + add) * 2 > c->hash.capacity) {
    cfb::PairHash nh;
    int rc = hash_alloc(&nh, pow2_at_least((exact + add) * 2), c->G, c->stream);
    if (rc) return rc;
    cfb::pair_hash_drain_kernel<<<148 * 4, 256, 0, c->stream>>>(c->hash, c->d_lay, c->d_lay, c->d_u64, nh, c->d_err, cfb::SlotTrans{}, nullptr, c->G);
    g_launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    hash_free(c->hash, c->G);
    c->hash = nh;
  }
  c->hash_upper += add;
  return CFB_OK;
}

int launch_remap_add(const Layout *d_dl, const Layout *d_sl, const Layout &sl, double *df, unsigned long long *du,
                     const double *sf, const unsigned long long *su, const cfb::PairHash &dhash, int *d_err,
                     cudaStream_t s, const cfb::SlotTrans &tr = cfb::SlotTrans{}, const int *group_map = nullptr) {
  const long long tot = (sl.F + sl.U) * sl.n_groups;
  const int blocks = (int)std::min<long long>((tot + 255) / 256, 148 * 8);
  cfb::remap_add_kernel<<<std::max(blocks, 1), 256, 0, s>>>(d_dl, d_sl, df, du, sf, su, dhash, d_err, tr, group_map);
  g_launches++;
  CU(cudaGetLastError());
  return CFB_OK;
}

long long dense_pair_entries(const Layout &L) {
  long long po = 0;
  for (int k = 0; k < L.m; k++)
    for (int l = k + 1; l < L.m; l++) po += (long long)L.dom[k] * L.dom[l];
  return po;
}

// Make the context's categorical domain cover [lo, hi] per column, re-laying out the state if
// it has to grow (stream-ordered).  A grown domain may switch the pair counts from dense
// tables to the hash table.
int ensure_domain(cfb_ctx *c, const int *lo, const int *hi, bool exact = false, const cfb::SlotTrans *rekey = nullptr) {
  if (c->m == 0) return CFB_OK;
  int nlo[cfb::kMaxCat], nhi[cfb::kMaxCat];
  bool grow = !c->lay.has_domain;
  for (int k = 0; k < c->m; k++) {
    if (exact || (rekey && rekey->col[k])) {
      // an untouched (fresh or recycled) context takes exactly the declared domain; a column that
      // is being re-keyed (dense -> dictionary codes) takes its new code range
      nlo[k] = lo[k];
      nhi[k] = hi[k];
      if (!c->lay.has_domain || c->lay.lo[k] != lo[k] || (long long)c->lay.lo[k] + c->lay.dom[k] - 1 != hi[k]) grow = true;
      if (rekey && rekey->col[k]) grow = true;
    } else if (c->lay.has_domain) {
      const int clo = c->lay.lo[k], chi = (int)((long long)c->lay.lo[k] + c->lay.dom[k] - 1);
      nlo[k] = std::min(clo, lo[k]);
      nhi[k] = std::max(chi, hi[k]);
      if (nlo[k] != clo || nhi[k] != chi) grow = true;
    } else {
      nlo[k] = lo[k];
      nhi[k] = hi[k];
    }
    if ((long long)nhi[k] - nlo[k] + 1 > (1ll << cfb::kPairSlotBits))
      return fail(CFB_ERR_DOMAIN, "categorical column %d spans [%d,%d]: key range too large", k, nlo[k], nhi[k]);
  }
  if (!grow) return CFB_OK;
  Layout nl;
  build_layout(nl, c->kind, c->n, c->m, c->G, nlo, nhi);
  double *nf = nullptr;
  unsigned long long *nu = nullptr;
  int rc = alloc_state(nl, &nf, &nu, c->stream);
  if (rc) return rc;
  cfb::PairHash nh{};
  if (nl.pairs_hashed) {
    // room for what the old state can hold: its hash entries, or its dense non-zeros (<= table size)
    unsigned long long carry = c->lay.pairs_hashed ? c->hash_upper : (unsigned long long)dense_pair_entries(c->lay);
    carry = std::min<unsigned long long>(carry, exact ? 0 : carry);
    rc = hash_alloc(&nh, pow2_at_least(2 * carry + 2), c->G, c->stream);
    if (rc) return rc;
  }
  Layout *d_nl = nullptr;
  CU(cudaMalloc(&d_nl, sizeof(Layout)));
  CU(cudaMemcpyAsync(d_nl, &nl, sizeof(Layout), cudaMemcpyHostToDevice, c->stream));
  // carry the old contents over; an old state without a domain has only its numeric part
  // and N, which the remap handles because its tables are empty
  if (!exact) {
    const cfb::SlotTrans tr = rekey ? *rekey : cfb::SlotTrans{};
    rc = launch_remap_add(d_nl, c->d_lay, c->lay, nf, nu, c->d_f64, c->d_u64, nh, c->d_err, c->stream, tr);
    if (rc) return rc;
    if (c->lay.pairs_hashed) {
      cfb::pair_hash_drain_kernel<<<148 * 4, 256, 0, c->stream>>>(c->hash, d_nl, c->d_lay, nu, nh, c->d_err, tr, nullptr, c->G);
      g_launches++;
      CU(cudaGetLastError());
    }
  }
  CU(cudaStreamSynchronize(c->stream));  // &nl and the old arrays must outlive the copies
  cudaFree(c->d_f64);
  cudaFree(c->d_u64);
  cudaFree(c->d_lay);
  if (c->hash.capacity) hash_free(c->hash, c->G);
  c->d_f64 = nf;
  c->d_u64 = nu;
  c->d_lay = d_nl;
  c->lay = nl;
  c->hash = nh;
  if (!nl.pairs_hashed || exact) c->hash_upper = 0;
  else if (!c->hash_upper) c->hash_upper = (nh.capacity - 2) / 2;  // dense -> hashed: unknown non-zero count
  return CFB_OK;
}

// ---- key dictionaries ----------------------------------------------------------------
// A categorical column switches to dictionary codes when its key range gets wider than this
// (CFB_DICT_RANGE overrides; tests force it to 1).
long long dict_range_limit() {
  static const long long v = [] {
    const char *e = getenv("CFB_DICT_RANGE");
    return e ? atoll(e) : (1ll << 22);
  }();
  return v;
}

void dict_free(cfb_ctx::ColDict &cd) {
  cudaFree(cd.d.table);
  cudaFree(cd.d.keys_of_code);
  cudaFree(cd.d.n_codes);
  cd = cfb_ctx::ColDict{};
}

bool any_dict(const cfb_ctx *c) {
  for (int k = 0; k < c->m; k++)
    if (c->dict[k].on) return true;
  return false;
}

// Room for `add` more distinct keys in column k's dictionary (table load <= 1/2).
int dict_reserve(cfb_ctx *c, int k, unsigned long long add, cudaStream_t s) {
  auto &cd = c->dict[k];
  const unsigned long long need = (unsigned long long)cd.n_codes + add;
  if (!cd.d.n_codes) {
    CU(cudaMalloc(&cd.d.n_codes, sizeof(int)));
    CU(cudaMemsetAsync(cd.d.n_codes, 0, sizeof(int), s));
  }
  if (need > cd.code_cap) {
    const unsigned long long cap = pow2_at_least(need);
    int *nk = nullptr;
    CU(cudaMalloc(&nk, cap * sizeof(int)));
    if (cd.n_codes) CU(cudaMemcpyAsync(nk, cd.d.keys_of_code, (size_t)cd.n_codes * sizeof(int), cudaMemcpyDeviceToDevice, s));
    CU(cudaStreamSynchronize(s));
    cudaFree(cd.d.keys_of_code);
    cd.d.keys_of_code = nk;
    cd.code_cap = cap;
  }
  if (need * 2 > cd.d.capacity) {
    cfb::KeyDict nd = cd.d;
    nd.capacity = pow2_at_least(need * 2);
    if (nd.capacity > (1ull << 31)) return fail(CFB_ERR_DOMAIN, "categorical column %d has too many distinct keys", k);
    CU(cudaMalloc(&nd.table, nd.capacity * 8));
    cfb::dict_clear_kernel<<<148 * 4, 256, 0, s>>>(nd.table, nd.capacity);
    if (cd.d.capacity) cfb::dict_rehash_kernel<<<148 * 4, 256, 0, s>>>(cd.d, nd);
    g_launches += 2;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s));
    cudaFree(cd.d.table);
    cd.d = nd;
  }
  return CFB_OK;
}

int dict_read_counts(cfb_ctx *c, cudaStream_t s) {
  for (int k = 0; k < c->m; k++)
    if (c->dict[k].on) CU(cudaMemcpyAsync(&c->dict[k].n_codes, c->dict[k].d.n_codes, sizeof(int), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return CFB_OK;
}

// Switch column k from dense slots (key - lo) to dictionary codes, carrying its contents over.
int dict_enable(cfb_ctx *c, int k, cudaStream_t s) {
  auto &cd = c->dict[k];
  if (cd.on) return CFB_OK;
  CU(cudaStreamSynchronize(c->stream));
  if (s != c->stream) CU(cudaStreamSynchronize(s));
  cd.on = true;
  const bool has_data = c->lay.has_domain && c->lay.dom[k] > 0;
  int rc = dict_reserve(c, k, has_data ? (unsigned long long)c->lay.dom[k] : 1024, c->stream);
  if (rc) return rc;
  int lo[cfb::kMaxCat], hi[cfb::kMaxCat];
  for (int j = 0; j < c->m; j++) {
    lo[j] = c->lay.has_domain ? c->lay.lo[j] : 0;
    hi[j] = c->lay.has_domain ? (int)((long long)c->lay.lo[j] + c->lay.dom[j] - 1) : 0;
  }
  if (!has_data) {
    if (c->lay.has_domain) {  // (cannot happen: a domain implies dom >= 1) keep the layout consistent
      lo[k] = hi[k] = 0;
    }
    return CFB_OK;
  }
  // codes for the keys that occurred so far, then re-key the column's slots through a translation
  int *trans = nullptr;
  const long long dom = c->lay.dom[k];
  CU(cudaMalloc(&trans, dom * sizeof(int)));
  const unsigned long long *counts = c->d_u64 + 1 + c->lay.cat_off[k];
  const int blocks = (int)std::min<long long>((dom + 255) / 256, 148 * 4);
  cfb::dict_translate_kernel<<<blocks, 256, 0, c->stream>>>(cd.d, nullptr, c->lay.lo[k], counts, c->lay.U, c->G, dom, trans, 0);
  cfb::dict_translate_kernel<<<blocks, 256, 0, c->stream>>>(cd.d, nullptr, c->lay.lo[k], counts, c->lay.U, c->G, dom, trans, 1);
  g_launches += 2;
  CU(cudaGetLastError());
  rc = dict_read_counts(c, c->stream);
  if (rc) return rc;
  lo[k] = 0;
  hi[k] = std::max(cd.n_codes - 1, 0);
  cfb::SlotTrans tr{};
  tr.col[k] = trans;
  rc = ensure_domain(c, lo, hi, false, &tr);
  cudaFree(trans);
  return rc;
}

// Before a scan of `rows` rows whose categorical columns live at cat[] on the device and span
// [obs_lo, obs_hi]: switch wide columns to dictionaries, give new keys their codes, grow the
// domain, and return in eff[] the columns the scan kernels should read (codes for dictionary
// columns: rewritten in place when `in_place`, else into the context's remap buffer).
int prepare_cats(cfb_ctx *c, const int32_t *const *cat, unsigned long long rows, cudaStream_t s, const int *obs_lo,
                 const int *obs_hi, bool in_place, const int32_t **eff) {
  const int m = c->m;
  for (int k = 0; k < m; k++) eff[k] = cat[k];
  if (m == 0 || rows == 0) return CFB_OK;
  int lo[cfb::kMaxCat], hi[cfb::kMaxCat];
  for (int k = 0; k < m; k++) {
    lo[k] = obs_lo[k];
    hi[k] = obs_hi[k];
    if (!c->dict[k].on && !c->user_domain) {
      long long l = lo[k], h = hi[k];
      if (c->lay.has_domain) {
        l = std::min<long long>(l, c->lay.lo[k]);
        h = std::max<long long>(h, (long long)c->lay.lo[k] + c->lay.dom[k] - 1);
      }
      if (h - l + 1 > dict_range_limit()) {
        int rc = dict_enable(c, k, s);
        if (rc) return rc;
      }
    }
  }
  if (any_dict(c)) {
    size_t n_dict = 0;
    for (int k = 0; k < m; k++)
      if (c->dict[k].on) {
        int rc = dict_reserve(c, k, rows, s);
        if (rc) return rc;
        const int blocks = (int)std::min<unsigned long long>((rows + 255) / 256, (unsigned long long)dev_info(c->device).sms * 8);
        cfb::dict_insert_kernel<<<blocks, 256, 0, s>>>(c->dict[k].d, cat[k], rows);
        g_launches++;
        n_dict++;
      }
    CU(cudaGetLastError());
    int rc = dict_read_counts(c, s);
    if (rc) return rc;
    if (!in_place) {
      if ((size_t)rows * n_dict > c->remap_cap) {
        CU(cudaStreamSynchronize(s));
        cudaFree(c->d_remap);
        c->d_remap = nullptr;
        CU(cudaMalloc(&c->d_remap, (size_t)rows * n_dict * sizeof(int32_t)));
        c->remap_cap = (size_t)rows * n_dict;
      }
      c->remap_rows = rows;
    }
    size_t j = 0;
    for (int k = 0; k < m; k++)
      if (c->dict[k].on) {
        int32_t *dst = in_place ? const_cast<int32_t *>(cat[k]) : c->d_remap + j * c->remap_rows;
        const int blocks = (int)std::min<unsigned long long>((rows + 255) / 256, (unsigned long long)dev_info(c->device).sms * 8);
        cfb::dict_remap_kernel<<<blocks, 256, 0, s>>>(c->dict[k].d, cat[k], dst, rows);
        g_launches++;
        eff[k] = dst;
        lo[k] = 0;
        hi[k] = std::max(c->dict[k].n_codes - 1, 0);
        j++;
      }
    CU(cudaGetLastError());
  }
  if (!c->user_domain) return ensure_domain(c, lo, hi);
  return CFB_OK;
}

// Host keys of dictionary column k -> codes (in place), inserting new keys (lifted-triple path).
int dict_codes_for_host_keys(cfb_ctx *c, int k, std::vector<int32_t> &keys) {
  if (keys.empty()) return CFB_OK;
  cudaStream_t s = c->stream;
  int rc = dict_reserve(c, k, keys.size(), s);
  if (rc) return rc;
  int32_t *d = nullptr;
  CU(cudaMalloc(&d, keys.size() * sizeof(int32_t)));
  CU(cudaMemcpyAsync(d, keys.data(), keys.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  const int blocks = (int)std::min<size_t>((keys.size() + 255) / 256, 148 * 4);
  cfb::dict_insert_kernel<<<blocks, 256, 0, s>>>(c->dict[k].d, d, keys.size());
  cfb::dict_remap_kernel<<<blocks, 256, 0, s>>>(c->dict[k].d, d, d, keys.size());
  g_launches += 2;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(keys.data(), d, keys.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(&c->dict[k].n_codes, c->dict[k].d.n_codes, sizeof(int), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  cudaFree(d);
  return CFB_OK;
}

// ------------------------------------------------------------------ Gram dispatch
template <bool DIAG, int... Ns>
constexpr std::array<cudaError_t (*)(const cfb::GramLaunchParams &), sizeof...(Ns)> gram_table(
    std::integer_sequence<int, Ns...>) {
  return {{cfb::gram_launch<Ns + 1, DIAG>...}};
}
const auto kGramTriple = gram_table<false>(std::make_integer_sequence<int, CFB_MAX_NUM>{});
const auto kGramNb = gram_table<true>(std::make_integer_sequence<int, CFB_MAX_NUM>{});

template <int KIND, int... Ns>
constexpr std::array<cudaError_t (*)(const cfb::SlabLaunchParams &), sizeof...(Ns)> slab_table(
    std::integer_sequence<int, Ns...>) {
  return {{cfb::slab_launch<Ns, KIND>...}};
}
const auto kSlabTriple = slab_table<0>(std::make_integer_sequence<int, CFB_MAX_NUM + 1>{});
const auto kSlabNb = slab_table<1>(std::make_integer_sequence<int, CFB_MAX_NUM + 1>{});

template <int BITS, int... Ns>
constexpr std::array<cudaError_t (*)(const cfb::RoleLaunchParams &), sizeof...(Ns)> role_table(
    std::integer_sequence<int, Ns...>) {
  return {{cfb::role_launch<Ns, BITS>...}};
}
const auto kRole16 = role_table<16>(std::make_integer_sequence<int, CFB_MAX_NUM + 1>{});
const auto kRole32 = role_table<32>(std::make_integer_sequence<int, CFB_MAX_NUM + 1>{});

template <int... Ns>
constexpr std::array<cudaError_t (*)(const cfb::BucketLaunchParams &), sizeof...(Ns)> bucket_table(std::integer_sequence<int, Ns...>) {
  return {{cfb::bucket_launch<Ns>...}};
}
const auto kBucket = bucket_table(std::make_integer_sequence<int, CFB_MAX_NUM + 1>{});

template <int... Ns>
constexpr std::array<cudaError_t (*)(const cfb::ChainLaunchParams &), sizeof...(Ns)> chain_table(std::integer_sequence<int, Ns...>) {
  return {{cfb::chain_launch<Ns>...}};
}
const auto kChain = chain_table(std::make_integer_sequence<int, CFB_MAX_NUM + 1>{});

int env_int(const char *name, int dflt) {
  const char *e = getenv(name);
  return e ? atoi(e) : dflt;
}

int launch_gram(cfb_ctx *c, const float *const *cols, unsigned long long rows, cudaStream_t s) {
  cfb::GramLaunchParams p{};
  p.cols = cols;
  p.rows = rows;
  p.smem_optin = dev_info(c->device).smem_optin;
  p.max_grid = c->gram_grid;
  p.stages = env_int("CFB_GRAM_STAGES", 0);
  p.flush_tiles = env_int("CFB_GRAM_FLUSH_TILES", 0);
  p.partials = c->d_partials;
  p.state = c->d_f64;  // ungrouped: slot 0, [lin | quad] leads the f64 array
  p.ticket = c->d_ticket;
  p.count = c->d_u64;  // ungrouped: N of slot 0
  p.stream = s;
  p.device = c->device;
  const cudaError_t e = (c->kind == CFB_NB ? kGramNb : kGramTriple)[c->n - 1](p);
  g_launches++;
  if (e != cudaSuccess) return fail(CFB_ERR_CUDA, "Gram kernel launch (n=%d): %s", c->n, cudaGetErrorString(e));
  return CFB_OK;
}

// Categorical / GROUP BY scan through the per-CTA fp32 slabs.  Returns 1 if the slab would be
// too large (caller falls back to generic_scan_kernel), 0 on success, <0 on error.
constexpr long long kMaxSlabBytesPerCta = 8ll << 20, kMaxSlabBytesTotal = 512ll << 20;

// Per-CTA fp32 slabs: `floats` per CTA for `grid` CTAs, all zero.
int ensure_slab(cfb_ctx *c, long long floats, int grid, cudaStream_t s) {
  const long long bytes = floats * 4;
  if (floats != c->slab_floats || grid > c->slab_grid) {
    if (c->d_slab) {
      CU(cudaStreamSynchronize(c->stream));
      if (c->user_stream) CU(cudaStreamSynchronize(c->user_stream));
      cudaFree(c->d_slab);
      c->d_slab = nullptr;
    }
    const int alloc_grid = std::max(grid, c->slab_grid);
    CU(cudaMalloc(&c->d_slab, (size_t)std::max<long long>(16, bytes * alloc_grid)));
    CU(cudaMemsetAsync(c->d_slab, 0, (size_t)std::max<long long>(16, bytes * alloc_grid), s));
    c->slab_floats = floats;
    c->slab_grid = alloc_grid;
  }
  return CFB_OK;
}

// Pair tables of the current domains cut into shared-memory roles (role_kernels.cuh).  Tables are
// taken in (k,l) order, so the tables of one role share their first column as far as possible.
bool build_role_plan(const Layout &L, size_t budget_bytes, int bits, cfb::RolePlan *out) {
  memset(out, 0, sizeof(*out));
  out->bits = bits;
  const long long budget_words = (long long)budget_bytes / 4;
  if (budget_words <= 0) return false;
  int role = 0;
  long long used = 0;
  for (int k = 0; k < L.m; k++) {
    for (int l = k + 1; l < L.m; l++) {
      const long long cells = (long long)L.dom[k] * L.dom[l];
      const long long gwords = bits == 32 ? cells : (cells + 1) / 2;
      const long long words = gwords * L.n_groups;  // one sub-table per GROUP BY slot
      if (words > budget_words) return false;
      if (used + words > budget_words || out->n_tables[role] == cfb::kRoleMaxTables) {
        if (++role == cfb::kRoleMaxRoles) return false;
        used = 0;
      }
      cfb::RoleTable &t = out->tbl[role][out->n_tables[role]++];
      t.k = k;
      t.l = l;
      t.dom_l = L.dom[l];
      t.cells = (int)cells;
      t.gwords = (int)gwords;
      t.word_off = (int)used;
      if (L.pair_off[k * L.m + l] > INT_MAX) return false;
      t.state_off = (int)L.pair_off[k * L.m + l];
      used += words;
      out->words[role] = (int)used;
    }
  }
  out->n_roles = role + 1;
  return true;
}

// Categorical part through shared-memory pair tables + L2 vector reductions.  Returns 1 if this
// shape / scan does not qualify (caller uses the slab kernel), 0 on success, <0 on error.
// Per-key payload sums of every categorical column by tile-level bucketing in shared memory
// (bucket_kernels.cuh).  Returns 1 if the shape does not qualify, 0 on success, <0 on error.
int launch_bucket(cfb_ctx *c, const cfb::ScanCols &sc, unsigned long long rows, cudaStream_t s) {
  if (c->kind != CFB_TRIPLE || c->m < 1 || getenv("CFB_NO_BUCKET")) return 1;
  if (c->lay.total_dom * c->G > cfb::kBucketMaxDom) return 1;
  const int P = cfb::pad4(1 + c->n), D = (int)c->lay.total_dom * c->G;  // buckets
  const long long budget = (long long)dev_info(c->device).smem_optin - 1024 - 8ll * D;
  int tile = (int)(budget / (4 * P + 2 * c->m)) / 256 * 256;  // any multiple of 4 rows works; 256 keeps the passes even
  tile = std::min(tile, 4096);
  if (tile < cfb::kBucketThreads) return 1;
  const unsigned long long n_tiles = (rows + tile - 1) / tile;
  const int grid = (int)std::min<unsigned long long>(dev_info(c->device).sms, n_tiles);
  int rc = ensure_slab(c, (long long)D * P, grid, s);
  if (rc) return rc;
  cfb::BucketLaunchParams p{};
  p.cols = sc;
  p.lay = &c->lay;
  p.rows = rows;
  p.tile_rows = tile;
  p.fold_tiles = std::max(1, 32768 / tile);  // an fp32 slab entry is folded into fp64 after at most ~32K rows of one CTA
  p.grid = grid;
  p.smem_max = dev_info(c->device).smem_optin - 1024;
  p.smem_bytes = cfb::bucket_smem_bytes(c->n, c->m, D, tile);
  p.slab = c->d_slab;
  p.f64 = c->d_f64;
  p.u64 = c->d_u64;
  p.err = c->d_err;
  p.stream = s;
  const cudaError_t e = kBucket[c->n](p);
  g_launches++;
  if (e != cudaSuccess) return fail(CFB_ERR_CUDA, "bucket kernel launch (n=%d): %s", c->n, cudaGetErrorString(e));
  return CFB_OK;
}

// Key counts + per-key sums through per-bucket linked lists in shared memory (chain_kernels.cuh); with `packed` the
// kernel also writes the validated one-byte slots of every row for pair_packed_kernel.  Returns 1 if the shape does
// not qualify, 0 on success, <0 on error.
int launch_chain(cfb_ctx *c, const cfb::ScanCols &sc, unsigned long long rows, unsigned char *packed, unsigned long long packed_stride,
                 cudaStream_t s) {
  if (c->kind != CFB_TRIPLE || c->m < 1 || getenv("CFB_NO_CHAIN") || getenv("CFB_NO_BUCKET")) return 1;
  const long long D = c->lay.total_dom * c->G;
  if (D > cfb::kChainMaxHeads || D < 1) return 1;
  const int QT = std::max(1, cfb::chain_quads(c->n));
  int sub_shift = 0;
  while ((D << (sub_shift + 1)) <= cfb::kChainMaxHeads && (D << sub_shift) * QT < 2048) sub_shift++;
  if (const char *e = getenv("CFB_CHAIN_SUB_SHIFT")) sub_shift = std::max(0, std::min(atoi(e), 12));
  while ((D << sub_shift) > cfb::kChainMaxHeads) sub_shift--;
  const int heads = (int)(D << sub_shift);
  // the skew plan (per-bucket sub-list counts) wants room for at least 2048 heads
  const bool adaptive = !getenv("CFB_CHAIN_NO_ADAPT") && c->n > 0;
  const int head_cap = adaptive ? std::min<int>(cfb::kChainMaxHeads, std::max(heads, 2048)) : heads;
  // kChainCtasPerSm CTAs share an SM: each gets its share of the SM's shared memory (1 KB per CTA is the system's)
  const long long cta_smem = std::min<long long>(dev_info(c->device).smem_optin - 1024,
                                                 (long long)dev_info(c->device).smem_sm / cfb::kChainCtasPerSm - 1024 - 512);
  const long long budget = cta_smem - (long long)cfb::chain_fixed_smem_bytes(head_cap, (int)c->lay.total_dom, (int)D);
  const int per_row = (c->n ? cfb::chain_quad_stride(c->n) * 16 : 0) + 2 * c->m;
  int tile = (int)std::min<long long>(budget / per_row, 32768);
  tile = tile >= 2 * cfb::kChainThreads ? tile / cfb::kChainThreads * cfb::kChainThreads : tile / 256 * 256;
  if (const char *e = getenv("CFB_CHAIN_TILE")) tile = std::min(tile, std::max(256, atoi(e) / 32 * 32));
  if (tile < 512) return 1;
  const unsigned long long n_tiles = (rows + tile - 1) / tile;
  const int grid = (int)std::min<unsigned long long>((unsigned long long)dev_info(c->device).sms * cfb::kChainCtasPerSm, n_tiles);
  const long long slab_floats = std::max<long long>(4, D * 4 * cfb::chain_quads(c->n));
  if (slab_floats != c->chain_slab_floats || grid > c->chain_slab_grid) {
    if (c->d_chain_slab) {
      CU(cudaStreamSynchronize(c->stream));
      if (c->user_stream) CU(cudaStreamSynchronize(c->user_stream));
      cudaFree(c->d_chain_slab);
      c->d_chain_slab = nullptr;
    }
    const int alloc_grid = std::max(grid, c->chain_slab_grid);
    CU(cudaMalloc(&c->d_chain_slab, (size_t)slab_floats * alloc_grid * sizeof(float)));
    CU(cudaMemsetAsync(c->d_chain_slab, 0, (size_t)slab_floats * alloc_grid * sizeof(float), s));
    c->chain_slab_floats = slab_floats;
    c->chain_slab_grid = alloc_grid;
  }
  if (D != c->cnt_slab_words || grid > c->cnt_slab_grid) {
    if (c->d_cnt_slab) {
      CU(cudaStreamSynchronize(c->stream));
      if (c->user_stream) CU(cudaStreamSynchronize(c->user_stream));
      cudaFree(c->d_cnt_slab);
      c->d_cnt_slab = nullptr;
    }
    const int alloc_grid = std::max(grid, c->cnt_slab_grid);
    CU(cudaMalloc(&c->d_cnt_slab, (size_t)D * alloc_grid * sizeof(unsigned)));
    CU(cudaMemsetAsync(c->d_cnt_slab, 0, (size_t)D * alloc_grid * sizeof(unsigned), s));
    c->cnt_slab_words = D;
    c->cnt_slab_grid = alloc_grid;
  }
  cfb::ChainLaunchParams p{};
  p.cols = sc;
  p.lay = &c->lay;
  p.rows = rows;
  p.tile_rows = tile;
  p.fold_tiles = std::max(1, 32768 / tile);  // an fp32 slab entry is folded into fp64 after at most ~32K rows of one CTA
  p.sub_shift = sub_shift;
  p.grid = grid;
  p.smem_max = dev_info(c->device).smem_optin - 1024;
  p.smem_bytes = cfb::chain_smem_bytes(c->n, c->m, head_cap, (int)c->lay.total_dom, (int)D, tile);
  p.head_cap = head_cap;
  p.adaptive = adaptive ? 1 : 0;
  p.slab = c->d_chain_slab;
  p.cnt_slab = c->d_cnt_slab;
  p.f64 = c->d_f64;
  p.u64 = c->d_u64;
  p.err = c->d_err;
  p.packed = packed;
  p.packed_stride = packed_stride;
  p.stream = s;
  const cudaError_t e = kChain[c->n](p);
  g_launches++;
  if (e != cudaSuccess) return fail(CFB_ERR_CUDA, "chain kernel launch (n=%d): %s", c->n, cudaGetErrorString(e));
  return CFB_OK;
}

// Room for `cols` packed one-byte columns of `rows` rows (column stride returned in *stride).
// Large packed-slot scratch (hundreds of MB for a 32 M-row slice) is not kept by parked contexts; it waits here for
// the next context instead of going back to the driver: a cudaMalloc of that size inside a scan costs milliseconds
// and now and then far more.  Buffers arrive with their context's streams drained.
struct PackedPool {
  std::mutex mu;
  struct Buf {
    int device;
    unsigned char *p;
    size_t cap;
  };
  std::vector<Buf> idle;
  static constexpr size_t kMaxIdle = 2;
  unsigned char *take(int device, size_t need, size_t *cap) {
    std::lock_guard<std::mutex> g(mu);
    for (size_t i = 0; i < idle.size(); i++)
      if (idle[i].device == device && idle[i].cap >= need) {
        unsigned char *p = idle[i].p;
        *cap = idle[i].cap;
        idle[i] = idle.back();
        idle.pop_back();
        return p;
      }
    return nullptr;
  }
  void give(int device, unsigned char *p, size_t cap) {
    std::lock_guard<std::mutex> g(mu);
    if (idle.size() >= kMaxIdle) {  // keep the larger ones
      size_t smallest = 0;
      for (size_t i = 1; i < idle.size(); i++)
        if (idle[i].cap < idle[smallest].cap) smallest = i;
      if (idle[smallest].cap >= cap) {
        cudaFree(p);
        return;
      }
      cudaFree(idle[smallest].p);
      idle[smallest] = idle.back();
      idle.pop_back();
    }
    idle.push_back({device, p, cap});
  }
};
PackedPool g_packed_pool;

int ensure_packed(cfb_ctx *c, int cols, unsigned long long rows, unsigned long long *stride) {
  *stride = (rows + 255) & ~255ull;
  const size_t need = (size_t)cols * *stride;
  if (need > c->packed_cap) {
    if (c->d_packed) {
      CU(cudaStreamSynchronize(c->stream));
      if (c->user_stream) CU(cudaStreamSynchronize(c->user_stream));
      g_packed_pool.give(c->device, c->d_packed, c->packed_cap);
      c->d_packed = nullptr;
      c->packed_cap = 0;
    }
    size_t cap = 0;
    if (unsigned char *p = g_packed_pool.take(c->device, need, &cap)) {
      c->d_packed = p;
      c->packed_cap = cap;
      return CFB_OK;
    }
    CU(cudaMalloc(&c->d_packed, need));
    c->packed_cap = need;
  }
  return CFB_OK;
}

// Key counts of the Naive-Bayes ring through a shared-memory histogram.  1 = shape does not qualify.
int launch_key_count(cfb_ctx *c, const cfb::ScanCols &sc, unsigned long long rows, cudaStream_t s) {
  if (c->kind != CFB_NB || c->m < 1 || getenv("CFB_NO_BUCKET")) return 1;
  const long long D = c->lay.total_dom * c->G;
  const int smem_max = dev_info(c->device).smem_optin - 1024;
  if (D * 4 > smem_max) return 1;
  cfb::KeyCountArgs a{};
  a.cols = sc;
  a.n_rows = rows;
  a.m = c->m;
  a.total_dom = (int)c->lay.total_dom;
  a.n_groups = c->G;
  a.U = c->lay.U;
  for (int k = 0; k < cfb::kMaxCat; k++) {
    a.lo[k] = c->lay.lo[k];
    a.dom[k] = c->lay.dom[k];
  }
  for (int k = 0; k <= cfb::kMaxCat; k++) a.cat_off[k] = (int)c->lay.cat_off[k];
  a.u64 = c->d_u64;
  a.err = c->d_err;
  static std::once_flag once[64];
  cudaError_t attr_err = cudaSuccess;
  std::call_once(once[c->device & 63], [&] {
    attr_err = cudaFuncSetAttribute(cfb::key_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
  });
  if (attr_err != cudaSuccess) return fail(CFB_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  const int per_sm = (int)std::max<long long>(1, std::min<long long>(2, smem_max / std::max<long long>(D * 4, 1)));
  const int grid = (int)std::max<unsigned long long>(1, std::min<unsigned long long>((unsigned long long)dev_info(c->device).sms * per_sm,
                                                                                 (rows + cfb::kBucketThreads - 1) / cfb::kBucketThreads));
  cfb::key_count_kernel<<<grid, cfb::kBucketThreads, (size_t)D * 4, s>>>(a);
  g_launches++;
  CU(cudaGetLastError());
  return CFB_OK;
}

// Is the role kernel usable for this scan?  Builds / refreshes the plan.  1 = no, 0 = yes.
int role_prepare(cfb_ctx *c, const cfb::ScanCols &sc, unsigned long long rows) {
  if (c->kind != CFB_TRIPLE || c->m < 2 || c->lay.pairs_hashed || getenv("CFB_NO_ROLE")) return 1;
  (void)rows;
  for (int k = 0; k < c->m; k++)
    if ((uintptr_t)sc.cat[k] & 15) return 1;  // the kernel reads 4 rows of a column per 128-bit load
  if ((uintptr_t)sc.group & 15) return 1;
  if (c->role_state == 1 && memcmp(c->role_dom, c->lay.dom, sizeof(int) * c->m) != 0) c->role_state = 0;
  if (c->role_state == 0) {
    c->role_state = -1;
    memcpy(c->role_dom, c->lay.dom, sizeof(c->role_dom));
    const int force_bits = env_int("CFB_ROLE_BITS", 0);
    const size_t budget = (size_t)dev_info(c->device).smem_optin - 1024;
    cfb::RolePlan p32, p16;
    const bool ok32 = force_bits != 16 && build_role_plan(c->lay, budget, 32, &p32);
    const bool ok16 = force_bits != 32 && build_role_plan(c->lay, budget, 16, &p16);
    // 32-bit cells fold once per scan instead of every 65 K rows, which measured faster than the fewer
    // passes over the key columns that 16-bit cells buy (C3: 9 roles 4.4 G rows/s, 5 roles 4.2 G rows/s)
    const int max_roles = std::max(1, dev_info(c->device).sms / 8);
    const cfb::RolePlan *pick = nullptr;
    if (ok32 && p32.n_roles <= max_roles) pick = &p32;
    else if (ok16) pick = &p16;
    if (pick && pick->n_roles <= max_roles) {
      if (!c->role_plan) c->role_plan = new cfb::RolePlan;
      *c->role_plan = *pick;
      int words = 0;
      for (int r = 0; r < pick->n_roles; r++) words = std::max(words, pick->words[r]);
      c->role_roles = pick->n_roles;
      c->role_bits = pick->bits;
      c->role_smem = (size_t)words * 4;
      c->role_state = 1;
    }
  }
  return c->role_state == 1 ? 0 : 1;
}

// Pair counts through shared-memory tables (+ the per-key payloads as L2 vector reductions when
// `do_sums`); call role_prepare first.
int launch_role(cfb_ctx *c, const cfb::ScanCols &sc, unsigned long long rows, bool do_sums, cudaStream_t s) {
  const int n_reps = std::max(1, dev_info(c->device).sms / c->role_roles);
  long long chunk = (long long)((rows + n_reps - 1) / n_reps);
  chunk = std::min<long long>(cfb::kRoleMaxChunkRows, std::max<long long>(1024, (chunk + 1023) / 1024 * 1024));
  const cfb::SlabShape sh = cfb::slab_shape(c->lay, 0);
  // several fp32 slabs per CTA while they stay small (they live in L2): thins out same-address reductions
  const int grid = c->role_roles * n_reps;
  int n_sub = std::max(1, env_int("CFB_ROLE_SUBSLABS", 4));
  while (n_sub > 1 && sh.floats * 4 * grid * n_sub > (48ll << 20)) n_sub /= 2;
  if (!do_sums) n_sub = 1;
  else {
    int rc = ensure_slab(c, sh.floats, grid * n_sub, s);
    if (rc) return rc;
  }
  cfb::RoleLaunchParams p{};
  p.cols = sc;
  p.lay = &c->lay;
  p.plan = c->role_plan;
  p.rows = rows;
  p.chunk_rows = (int)chunk;
  p.pair_fold_chunks = c->role_bits == 16 ? std::max(1, cfb::kRoleFoldRows16 / (int)chunk) : (1 << 30) / (int)chunk;
  p.n_roles = c->role_roles;
  p.n_reps = n_reps;
  p.smem_max = dev_info(c->device).smem_optin - 1024;
  p.smem_bytes = c->role_smem;
  p.skip = env_int("CFB_ROLE_DEBUG", 0) | (do_sums ? 0 : 2);
  p.n_sub = n_sub;
  p.slab = c->d_slab;
  p.f64 = c->d_f64;
  p.u64 = c->d_u64;
  p.err = c->d_err;
  p.stream = s;
  const cudaError_t e = (c->role_bits == 16 ? kRole16 : kRole32)[c->n](p);
  g_launches++;
  if (e != cudaSuccess) return fail(CFB_ERR_CUDA, "role kernel launch (n=%d): %s", c->n, cudaGetErrorString(e));
  return CFB_OK;
}

// Pair counts from the packed slots chain_sum_kernel wrote (call role_prepare first: same plan, roles and replicas).
int launch_pair_packed(cfb_ctx *c, unsigned long long rows, const unsigned char *packed, unsigned long long stride, bool grouped,
                       cudaStream_t s) {
  const int n_reps = std::max(1, dev_info(c->device).sms / c->role_roles);
  const unsigned long long rows16 = grouped ? (rows + 15) & ~15ull : rows & ~15ull;  // rows done in 16-row groups
  const int step = 16 * cfb::kRoleThreads;  // rows of one pass of the CTA over a chunk
  // 16-bit cells are folded after every chunk (<= 65024 rows); 32-bit cells once, at the end of the scan
  const int max_steps = c->role_bits == 16 ? 3 : 4;
  long long chunk = (long long)((rows16 + n_reps - 1) / n_reps);
  chunk = std::min<long long>((long long)max_steps * step, std::max<long long>(step, (chunk + step - 1) / step * step));
  cfb::PackedPairArgs a{};
  a.packed = packed;
  a.stride = stride;
  a.n_rows = grouped ? rows16 : rows;
  a.chunk_rows = (int)chunk;
  a.pair_fold_chunks = c->role_bits == 16 ? 1 : 1 << 30;
  a.n_reps = n_reps;
  a.m = c->m;
  a.n_groups = c->G;
  a.U = c->lay.U;
  a.pair_base = c->lay.pair_base;
  a.u64 = c->d_u64;
  a.plan = *c->role_plan;
  const int smem_max = dev_info(c->device).smem_optin - 1024;
  auto kern = c->role_bits == 16 ? (grouped ? cfb::pair_packed_kernel<16, true> : cfb::pair_packed_kernel<16, false>)
                                 : (grouped ? cfb::pair_packed_kernel<32, true> : cfb::pair_packed_kernel<32, false>);
  CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
  kern<<<c->role_roles * n_reps, cfb::kRoleThreads, c->role_smem, s>>>(a);
  g_launches++;
  CU(cudaGetLastError());
  return CFB_OK;
}

int launch_slab(cfb_ctx *c, const cfb::ScanCols &sc, unsigned long long rows, int do_numeric, cudaStream_t s) {
  const cfb::SlabShape sh = cfb::slab_shape(c->lay, do_numeric);
  const long long bytes = sh.floats * 4;
  if (bytes > kMaxSlabBytesPerCta || getenv("CFB_NO_SLAB")) return 1;
  int grid = dev_info(c->device).sms * 2;
  grid = (int)std::min<long long>(grid, std::max<long long>(1, kMaxSlabBytesTotal / std::max<long long>(bytes, 1)));
  grid = (int)std::min<unsigned long long>(grid, (rows + cfb::kSlabTile - 1) / cfb::kSlabTile);
  grid = std::max(grid, 1);
  int erc = ensure_slab(c, sh.floats, grid, s);
  if (erc) return erc;
  cfb::SlabLaunchParams p{};
  p.cols = sc;
  p.d_lay = c->d_lay;
  p.rows = rows;
  p.do_numeric = do_numeric;
  // an fp32 slab entry is folded into fp64 after at most ~32K rows of one CTA
  p.flush_tiles = std::max(1, env_int("CFB_SLAB_FLUSH_TILES", 32));
  p.slab = c->d_slab;
  p.f64 = c->d_f64;
  p.u64 = c->d_u64;
  p.err = c->d_err;
  p.hash = c->hash;
  p.grid = grid;
  p.stream = s;
  const cudaError_t e = (c->kind == CFB_NB ? kSlabNb : kSlabTriple)[c->n](p);
  g_launches++;
  if (e != cudaSuccess) return fail(CFB_ERR_CUDA, "slab kernel launch (n=%d): %s", c->n, cudaGetErrorString(e));
  return CFB_OK;
}

// GROUP BY / filtered numeric part through the warp-private shared-memory tables.  Returns 1 if
// the tables do not fit (caller lets the slab kernel do the numeric part), 0 on success.
template <int E>
int launch_group_e(cfb_ctx *c, const cfb::GroupArgs &a, size_t smem, int grid, cudaStream_t s) {
  auto kern = cfb::group_scan_kernel<E>;
  static std::once_flag once[64];
  cudaError_t attr_err = cudaSuccess;
  std::call_once(once[c->device & 63], [&] {
    attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dev_info(c->device).smem_optin - 1024);
  });
  if (attr_err != cudaSuccess) return fail(CFB_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  kern<<<grid, cfb::kGroupThreads, smem, s>>>(a);
  g_launches++;
  CU(cudaGetLastError());
  return CFB_OK;
}

// GROUP BY / filtered numeric part with the rows bucketed by slot and the accumulators in registers
// (slot_gram_kernel.cuh).  Returns 1 if the shape does not qualify (caller uses group_scan_kernel).
template <int E, int TPW>
int launch_slot_gram_e(cfb_ctx *c, const cfb::SlotGramArgs &a, size_t smem, int grid, cudaStream_t s) {
  auto kern = cfb::slot_gram_kernel<E, TPW>;
  static std::once_flag once[64];
  cudaError_t attr_err = cudaSuccess;
  std::call_once(once[c->device & 63], [&] {
    attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dev_info(c->device).smem_optin - 1024);
  });
  if (attr_err != cudaSuccess) return fail(CFB_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  kern<<<grid, cfb::kSlotThreads, smem, s>>>(a);
  g_launches++;
  CU(cudaGetLastError());
  return CFB_OK;
}

// Few slots: bulk-copy fetch, atomic-rank sort and 4x4 register blocks + FFMA2 over the slot-sorted tile
// (slot_block_kernel.cuh).  1 = the shape does not qualify (more (slot, block) lane tasks than a CTA has lanes).
template <bool NB>
int launch_slot_block_k(cfb_ctx *c, const cfb::SlotBlockArgs &a, size_t smem, int grid, cudaStream_t s) {
  auto kern = cfb::slot_block_kernel<NB>;
  static std::once_flag once[64];
  cudaError_t attr_err = cudaSuccess;
  std::call_once(once[c->device & 63], [&] {
    attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dev_info(c->device).smem_optin - 1024);
  });
  if (attr_err != cudaSuccess) return fail(CFB_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  kern<<<grid, cfb::kSlotThreads, smem, s>>>(a);
  g_launches++;
  CU(cudaGetLastError());
  return CFB_OK;
}

int launch_slot_block(cfb_ctx *c, const cfb::ScanCols &sc, unsigned long long rows, cudaStream_t s, bool *keys_done) {
  *keys_done = false;
  if (c->n > 31 || getenv("CFB_NO_SLOT_BLOCK")) return 1;  // (n + 1 columns are fetched by one warp)
  const bool nb_ring = c->kind == CFB_NB;
  // Naive-Bayes ring: the key counts ride along when their columns fit the fetch warp and the histogram stays small
  int fuse_m = 0;
  if (nb_ring && c->m > 0 && c->n + 1 + c->m <= 32 && c->lay.total_dom * c->G * 4 <= 32768 && !getenv("CFB_NO_NB_FUSE")) {
    fuse_m = c->m;
    for (int k = 0; k < c->m; k++)
      if ((uintptr_t)sc.cat[k] & 15) fuse_m = 0;
  }
  const int td = fuse_m ? (int)c->lay.total_dom : 0;
  const int nblk = cfb::slotb_blocks(c->n, nb_ring);
  if (c->G * nblk > cfb::kSlotThreads) return 1;
  if (!sc.group || ((uintptr_t)sc.group & 15)) return 1;  // the tile is fetched with 16-byte bulk copies
  int steps = 0, per_sm = 1;
  for (int want : {2, 1}) {
    const size_t budget = want == 1 ? (size_t)dev_info(c->device).smem_optin - 1024
                                    : (size_t)(dev_info(c->device).smem_sm - 1024 * want) / want - 512;
    int st = cfb::kSlotMaxSteps;
    while (st >= 1 && cfb::slotb_smem_bytes(c->n, c->G, st, fuse_m, td) > budget) st--;
    if (st >= (want == 1 ? 1 : 2)) {
      steps = st;
      per_sm = want;
      break;
    }
  }
  if (const char *e = getenv("CFB_SLOT_STEPS")) steps = std::max(1, std::min(steps, atoi(e)));
  if (steps < 1) return 1;
  cfb::SlotBlockArgs a{};
  a.cols = sc;
  a.n_rows = rows;
  a.n = c->n;
  a.n_groups = c->G;
  a.steps = steps;
  a.splits = std::max(1, cfb::kSlotThreads / (c->G * nblk));
  const int tile = steps * cfb::kSlotThreads;
  a.fold_tiles = std::max(1, 32768 / tile);  // an fp32 accumulator is folded into fp64 after at most ~32K rows of one CTA
  a.F = c->lay.F;
  a.U = c->lay.U;
  a.f64 = c->d_f64;
  a.u64 = c->d_u64;
  a.err = c->d_err;
  a.m = fuse_m;
  a.total_dom = td;
  for (int k = 0; k < cfb::kMaxCat; k++) {
    a.lo[k] = c->lay.lo[k];
    a.dom[k] = c->lay.dom[k];
  }
  for (int k = 0; k <= cfb::kMaxCat; k++) a.cat_off[k] = (int)c->lay.cat_off[k];
  const size_t smem = cfb::slotb_smem_bytes(c->n, c->G, steps, fuse_m, td);
  const int grid = (int)std::min<unsigned long long>((unsigned long long)dev_info(c->device).sms * per_sm, (rows + tile - 1) / tile);
  *keys_done = fuse_m > 0;
  return nb_ring ? launch_slot_block_k<true>(c, a, smem, grid, s) : launch_slot_block_k<false>(c, a, smem, grid, s);
}

int launch_slot_gram(cfb_ctx *c, const cfb::ScanCols &sc, unsigned long long rows, cudaStream_t s, bool *keys_done) {
  *keys_done = false;
  if (getenv("CFB_NO_SLOT_GRAM") || c->n < 1 || c->G > cfb::kSlotMaxGroups) return 1;
  if (rows < (unsigned long long)std::max(1, env_int("CFB_SLOT_MIN_ROWS", 8192))) return 1;
  {
    const int rc = launch_slot_block(c, sc, rows, s, keys_done);
    if (rc <= 0) return rc;
  }
  // 2x2 blocks of a slot over `parts` warps, E per lane; the remaining warps split the slot's rows.  A warp carries
  // two tasks when E == 1 (8 accumulators per block and task).
  const int V = cfb::slot_blocks(c->n, c->kind);
  int parts = 0, E = 0, tpw = 0;
  for (int t : {2, 1}) {
    const int parts_max = t * cfb::kSlotWarps / c->G, e_max = t == 2 ? 1 : 4;
    if (parts_max < 1) continue;
    int p = 1;
    while (p < parts_max && (V + 32 * p - 1) / (32 * p) > e_max) p++;
    const int e = (V + 32 * p - 1) / (32 * p);
    if (e > e_max) continue;
    parts = p;
    E = e == 3 ? 4 : e;
    tpw = t;
    break;
  }
  if (!parts) return 1;
  const int smem_max = dev_info(c->device).smem_optin - 1024;
  // as many CTAs per SM (up to 4) as still leave a tile of >= 2 rows per thread
  int steps = 0, per_sm = 1;
  for (int want : {4, 3, 2, 1}) {
    if (want > 2 && E > 1) continue;  // those instantiations take 128 registers (launch bounds)
    const size_t budget = want == 1 ? (size_t)smem_max : (size_t)(dev_info(c->device).smem_sm - 1024 * want) / want - 512;
    int st = cfb::kSlotMaxSteps;
    while (st >= 1 && cfb::slot_smem_bytes(c->n, c->G, st) > budget) st--;
    if (st >= (want == 1 ? 1 : 2)) {
      steps = st;
      per_sm = want;
      break;
    }
  }
  if (steps < 1) return 1;
  cfb::SlotGramArgs a{};
  a.cols = sc;
  a.n_rows = rows;
  a.n = c->n;
  a.kind = c->kind;
  a.n_groups = c->G;
  a.steps = steps;
  a.parts = parts;
  a.splits = std::max(1, tpw * cfb::kSlotWarps / (c->G * parts));
  const int tile = steps * cfb::kSlotThreads;
  a.fold_tiles = std::max(1, 32768 / tile);  // an fp32 accumulator is folded into fp64 after at most ~32K rows of one CTA
  a.F = c->lay.F;
  a.U = c->lay.U;
  a.f64 = c->d_f64;
  a.u64 = c->d_u64;
  a.err = c->d_err;
  const size_t smem = cfb::slot_smem_bytes(c->n, c->G, steps);
  const int grid = (int)std::min<unsigned long long>((unsigned long long)dev_info(c->device).sms * per_sm, (rows + tile - 1) / tile);
  switch (E) {
    case 1: return tpw == 2 ? launch_slot_gram_e<1, 2>(c, a, smem, grid, s) : launch_slot_gram_e<1, 1>(c, a, smem, grid, s);
    case 2: return launch_slot_gram_e<2, 1>(c, a, smem, grid, s);
    default: return launch_slot_gram_e<4, 1>(c, a, smem, grid, s);
  }
}

int launch_group(cfb_ctx *c, const cfb::ScanCols &sc, unsigned long long rows, cudaStream_t s, bool *keys_done) {
  {
    const int rc = launch_slot_gram(c, sc, rows, s, keys_done);
    if (rc <= 0) return rc;
  }
  if (getenv("CFB_NO_GROUP_KERNEL")) return 1;
  const int V = cfb::group_entries(c->n, c->kind);
  const int need = (V + 31) / 32;
  static const int kE[] = {1, 2, 3, 4, 6, 8, 12, 18};
  int E = 0;
  for (int e : kE)
    if (e >= need) {
      E = e;
      break;
    }
  if (!E) return 1;
  const size_t smem = cfb::group_smem_bytes(c->n, c->G, E);
  if (smem > (size_t)dev_info(c->device).smem_optin - 1024) return 1;
  cfb::GroupArgs a{};
  a.cols = sc;
  a.n_rows = rows;
  a.n = c->n;
  a.kind = c->kind;
  a.n_groups = c->G;
  a.flush_rows = std::max(256, env_int("CFB_GROUP_FLUSH_ROWS", 2048));
  a.F = c->lay.F;
  a.U = c->lay.U;
  a.f64 = c->d_f64;
  a.u64 = c->d_u64;
  a.err = c->d_err;
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, ((size_t)dev_info(c->device).smem_optin - 1024) / std::max<size_t>(smem, 1)));
  int grid = dev_info(c->device).sms * per_sm;
  grid = (int)std::max<unsigned long long>(1, std::min<unsigned long long>(grid, (rows + 32 * cfb::kGroupWarps - 1) / (32 * cfb::kGroupWarps)));
  switch (E) {
    case 1: return launch_group_e<1>(c, a, smem, grid, s);
    case 2: return launch_group_e<2>(c, a, smem, grid, s);
    case 3: return launch_group_e<3>(c, a, smem, grid, s);
    case 4: return launch_group_e<4>(c, a, smem, grid, s);
    case 6: return launch_group_e<6>(c, a, smem, grid, s);
    case 8: return launch_group_e<8>(c, a, smem, grid, s);
    case 12: return launch_group_e<12>(c, a, smem, grid, s);
    default: return launch_group_e<18>(c, a, smem, grid, s);
  }
}

// -------------------------------------------------------------------- device scan
int scan_device(cfb_ctx *c, const float *const *num, const int32_t *const *cat, const int32_t *group,
                unsigned long long rows, cudaStream_t s) {
  if (rows == 0) return CFB_OK;
  c->touched = true;
  for (int k = 0; k < c->n; k++)
    if (!num[k] || ((uintptr_t)num[k] & 15))
      return fail(CFB_ERR_INVALID, "numeric column %d must be a 16-byte aligned device pointer", k);
  for (int k = 0; k < c->m; k++)
    if (!cat[k]) return fail(CFB_ERR_INVALID, "categorical column %d is NULL", k);
  cfb::ScanCols sc{};
  for (int k = 0; k < c->n; k++) sc.num[k] = num[k];
  for (int k = 0; k < c->m; k++) sc.cat[k] = cat[k];
  sc.group = group;
  const bool grouped = group != nullptr;
  if (c->timed) CU(cudaEventRecord(c->ev0, s));
  if (!grouped) {
    if (c->n > 0) {
      int rc = launch_gram(c, num, rows, s);  // (N += rows goes along)
      if (rc) return rc;
    } else {
      cfb::add_rows_kernel<<<1, 32, 0, s>>>(c->d_u64, rows);
      g_launches++;
    }
  }
  int slab_numeric = grouped ? 1 : 0;
  bool keys_done = false;  // the Naive-Bayes key counts went along with the numeric part
  if (grouped) {
    const int rc = launch_group(c, sc, rows, s, &keys_done);  // N / lin / quad of every slot (and the row filter)
    if (rc < 0) return rc;
    if (rc == 0) slab_numeric = 0;
  }
  if ((grouped && slab_numeric) || (c->m > 0 && !keys_done)) {
    // With hashed pair counts the scan is cut into slices so that the table can be grown
    // (host-side, between launches) before it could fill up.
    const int npairs = c->kind == CFB_TRIPLE ? c->m * (c->m - 1) / 2 : 0;
    const bool hashed = c->lay.pairs_hashed && npairs > 0;
    unsigned long long slice = rows;
    if (hashed) slice = std::max<unsigned long long>(1ull << 18, ((1ull << 24) / npairs) & ~1023ull);
    else if (c->kind == CFB_TRIPLE && c->m >= 2)  // bounds the packed-slot scratch of the pair kernel ((m+1) bytes per row)
      slice = std::max<unsigned long long>(1ull << 20, (unsigned long long)env_int("CFB_CAT_SLICE_ROWS", 32 << 20) & ~16383ull);
    // When the table for the WHOLE call fits the budget it is reserved once, here: growing slice by slice costs a
    // cudaMalloc, a rehash of everything seen so far and a stream synchronisation per doubling.
    bool reserved_whole = false;
    if (hashed && rows > slice) {
      const unsigned long long budget = (unsigned long long)std::max(0, env_int("CFB_HASH_RESERVE_MB", 8192)) << 20;
      const unsigned long long whole = std::min<unsigned long long>(rows * npairs, (unsigned long long)dense_pair_entries(c->lay));
      if (pow2_at_least((c->hash_upper + whole) * 2) * (unsigned long long)c->G * 16 <= budget) {
        if (s != c->stream) CU(cudaStreamSynchronize(s));
        int rc = hash_reserve(c, whole);
        if (rc) return rc;
        reserved_whole = true;
      }
    }
    for (unsigned long long r0 = 0; r0 < rows; r0 += slice) {
      const unsigned long long cnt = std::min(slice, rows - r0);
      cfb::ScanCols part = sc;
      for (int k = 0; k < c->n; k++) part.num[k] = num[k] + r0;
      for (int k = 0; k < c->m; k++) part.cat[k] = cat[k] + r0;
      if (group) part.group = group + r0;
      if (hashed && !reserved_whole) {
        if (s != c->stream) CU(cudaStreamSynchronize(s));
        const unsigned long long worst = std::min<unsigned long long>(cnt * npairs, (unsigned long long)dense_pair_entries(c->lay));
        int rc = hash_reserve(c, worst);
        if (rc) return rc;
      }
      if (!slab_numeric && c->kind == CFB_NB && cnt >= (unsigned long long)std::max(1, env_int("CFB_ROLE_MIN_ROWS", 16384))) {
        const int rc = launch_key_count(c, part, cnt, s);  // the NB ring keeps key counts only
        if (rc < 0) return rc;
        if (rc == 0) continue;
      }
      if (!slab_numeric && cnt >= (unsigned long long)std::max(1, env_int("CFB_ROLE_MIN_ROWS", 16384))) {
        // dense small domains: pair counts in shared-memory tables, per-key sums by tile bucketing
        const bool pairs_here = role_prepare(c, part, cnt) == 0;
        if (pairs_here || (c->m == 1 && c->kind == CFB_TRIPLE)) {
          // second generation: per-bucket lists (no second ranking pass) + pair counts from packed one-byte slots
          bool pack = pairs_here && c->G <= 255 && !getenv("CFB_NO_PACKED");
          for (int k = 0; k < c->m && pack; k++) pack = c->lay.dom[k] <= 255;
          unsigned long long stride = 0;
          if (pack) {
            const int rc = ensure_packed(c, c->m + (group ? 2 : 0), cnt, &stride);
            if (rc) return rc;
          }
          const int ch = launch_chain(c, part, cnt, pack ? c->d_packed : nullptr, stride, s);
          if (ch < 0) return ch;
          if (ch == 0) {
            if (pairs_here) {
              const int rc = pack ? launch_pair_packed(c, cnt, c->d_packed, stride, group != nullptr, s)
                                  : launch_role(c, part, cnt, /*do_sums=*/false, s);
              if (rc < 0) return rc;
            }
            continue;
          }
          const int b = launch_bucket(c, part, cnt, s);
          if (b < 0) return b;
          if (pairs_here && (b == 0 || c->G == 1)) {  // (the role kernel's own payload path has single-slot slabs)
            const int rc = launch_role(c, part, cnt, b != 0, s);
            if (rc < 0) return rc;
            continue;
          }
          if (b == 0 && c->m == 1) continue;
          if (b == 0) return fail(CFB_ERR_CUDA, "internal: bucket sums without pair counts");
        }
      }
      int need_generic = launch_slab(c, part, cnt, slab_numeric, s);
      if (need_generic < 0) return need_generic;
      if (need_generic) {  // 1 = slab too large for this shape
        const int blocks = (int)std::min<unsigned long long>((cnt + 255) / 256, (unsigned long long)dev_info(c->device).sms * 8);
        cfb::generic_scan_kernel<<<std::max(blocks, 1), 256, 0, s>>>(part, c->d_lay, cnt, slab_numeric, c->d_f64,
                                                                     c->d_u64, c->d_err, c->hash);
        g_launches++;
      }
    }
  }
  CU(cudaGetLastError());
  if (c->timed) CU(cudaEventRecord(c->ev1, s));
  return CFB_OK;
}

int check_async_error(cfb_ctx *c) {
  int e = 0;
  CU(cudaMemcpyAsync(&e, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (e) {
    CU(cudaMemsetAsync(c->d_err, 0, sizeof(int), c->stream));
    if (e == 1) return fail(CFB_ERR_DOMAIN, "a categorical key lies outside the declared domain");
    if (e == 3) return fail(CFB_ERR_CUDA, "internal: the pair hash table filled up");
    return fail(CFB_ERR_INVALID, "a group slot lies outside [0, n_groups)");
  }
  return CFB_OK;
}

// ------------------------------------------------------------------------ staging
size_t stage_cols(const cfb_ctx *c) { return (size_t)c->n + c->m + 1; }

int ensure_staging(cfb_ctx *c) {
  if (c->tile_rows) return CFB_OK;
  for (int i = 0; i < 2; i++) {
    int rc = g_stage_pool.acquire(c->device, &c->st[i]);
    if (rc) return rc;
  }
  size_t rows = std::min(c->st[0].bytes, c->st[1].bytes) / (4 * stage_cols(c));
  if (const char *e = getenv("CFB_STAGE_ROWS")) rows = std::min<size_t>(rows, (size_t)std::max(1ll, atoll(e)));
  rows = std::max<size_t>(256, rows / 256 * 256);
  c->tile_rows = rows;
  c->fill = 0;
  for (int k = 0; k < c->m; k++) {
    c->st_lo[k] = INT_MAX;
    c->st_hi[k] = INT_MIN;
  }
  return CFB_OK;
}

// Ship the open tile to the device and reduce it there; the host moves on to the other tile.
int flush_tile(cfb_ctx *c) {
  if (!c->fill) return CFB_OK;
  Stage &st = c->st[c->cur];
  const size_t rows = c->fill, tr = c->tile_rows, ncol = stage_cols(c);
  _mm_sfence();  // the staging tile was filled with non-temporal stores
  if (rows == tr) {
    CU(cudaMemcpyAsync(st.d, st.h, tr * 4 * (ncol - (c->uses_group ? 0 : 1)), cudaMemcpyHostToDevice, c->stream));
  } else {
    for (size_t k = 0; k < ncol; k++) {
      if (k == ncol - 1 && !c->uses_group) break;
      CU(cudaMemcpyAsync(st.d + k * tr * 4, st.h + k * tr * 4, rows * 4, cudaMemcpyHostToDevice, c->stream));
    }
  }
  const float *num[CFB_MAX_NUM];
  const int32_t *cat[cfb::kMaxCat];
  for (int k = 0; k < c->n; k++) num[k] = (const float *)(st.d + (size_t)k * tr * 4);
  for (int k = 0; k < c->m; k++) cat[k] = (const int32_t *)(st.d + (size_t)(c->n + k) * tr * 4);
  const int32_t *grp = c->uses_group ? (const int32_t *)(st.d + (size_t)(c->n + c->m) * tr * 4) : nullptr;
  // domain growth / key dictionaries for this tile's keys (the tile is ours: codes are written in place)
  const int32_t *eff[cfb::kMaxCat];
  int rc = prepare_cats(c, cat, rows, c->stream, c->st_lo, c->st_hi, /*in_place=*/true, eff);
  if (rc) return rc;
  rc = scan_device(c, num, eff, grp, rows, c->stream);
  if (rc) return rc;
  CU(cudaEventRecord(st.done, c->stream));
  st.in_flight = true;
  c->cur ^= 1;
  c->fill = 0;
  for (int k = 0; k < c->m; k++) {
    c->st_lo[k] = INT_MAX;
    c->st_hi[k] = INT_MIN;
  }
  Stage &nx = c->st[c->cur];
  if (nx.in_flight) {
    CU(cudaEventSynchronize(nx.done));
    nx.in_flight = false;
  }
  return CFB_OK;
}

// Host -> staging-tile copies live in stage_copy.cpp (host compiler: AVX-512 / AVX2 / SSE2 non-temporal stores).
}  // namespace
extern "C" void cfb_stage_copy(void *dst, const void *src, size_t bytes);
extern "C" void cfb_stage_gather32(void *dst, const void *src, const uint32_t *sel, size_t count);
extern "C" void cfb_stage_minmax32(const int32_t *src, size_t count, int32_t *lo, int32_t *hi);
namespace {

template <class T>
inline void gather(T *dst, const T *src, const uint32_t *sel, size_t first, size_t cnt) {
  static_assert(sizeof(T) == 4, "staged columns are 4-byte values");
  if (!sel)
    cfb_stage_copy(dst, src + first, cnt * sizeof(T));
  else
    cfb_stage_gather32(dst, src, sel + first, cnt);
}

}  // namespace

// ------------------------------------------------------------------ model scores (MICE write-back)
struct cfb_model {
  int device = 0;
  int type = 0;  // 0: linear scores (linreg / LDA), 1: Gaussian naive Bayes, 2: QDA
  double *d_model = nullptr;
  int *d_map = nullptr;
  int *d_labels = nullptr;  // naive Bayes / QDA
  cfb::PredictArgs args{};
  cfb::NbArgs nb{};
  cfb::QdaArgs qda{};
  size_t smem = 0;
};

namespace {

int predict_launch(cfb_model *M, const float *const *num, const int32_t *const *cat, const int32_t *mask, size_t rows,
                   int mode, void *d_out, cudaStream_t s) {
  if (M->type != 0) {
    // naive Bayes / QDA: one row per thread, the model read from global memory, the class label out
    if (mode != CFB_PREDICT_LABEL) return fail(CFB_ERR_INVALID, "a naive-Bayes / QDA model predicts labels (CFB_PREDICT_LABEL)");
    const int n = M->type == 1 ? M->nb.n : M->qda.n, m = M->type == 1 ? M->nb.m : M->qda.m;
    cfb::ScanCols cols{};
    for (int k = 0; k < n; k++) {
      if (!num[k]) return fail(CFB_ERR_INVALID, "numeric column %d is NULL", k);
      cols.num[k] = num[k];
    }
    for (int k = 0; k < m; k++) {
      if (!cat[k]) return fail(CFB_ERR_INVALID, "categorical column %d is NULL", k);
      cols.cat[k] = cat[k];
    }
    cols.group = mask;
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((size_t)dev_info(M->device).sms * 8, (rows + cfb::kPredictThreads - 1) / cfb::kPredictThreads));
    if (M->type == 1) {
      cfb::NbArgs a = M->nb;
      a.cols = cols;
      a.n_rows = rows;
      a.out = (int *)d_out;
      cfb::predict_nb_kernel<<<grid, cfb::kPredictThreads, 0, s>>>(a);
    } else {
      cfb::QdaArgs a = M->qda;
      a.cols = cols;
      a.n_rows = rows;
      a.out = (int *)d_out;
      cfb::predict_qda_kernel<<<grid, cfb::kPredictThreads, 0, s>>>(a);
    }
    g_launches++;
    CU(cudaGetLastError());
    return CFB_OK;
  }
  if (mode != CFB_PREDICT_SCORE && mode != CFB_PREDICT_ARGMAX) return fail(CFB_ERR_INVALID, "unknown predict mode %d", mode);
  cfb::PredictArgs a = M->args;
  bool aligned = (((uintptr_t)d_out | (uintptr_t)mask) & 15) == 0;
  for (int k = 0; k < a.n; k++) {
    if (!num[k]) return fail(CFB_ERR_INVALID, "numeric column %d is NULL", k);
    a.cols.num[k] = num[k];
    aligned = aligned && ((uintptr_t)num[k] & 15) == 0;
  }
  for (int k = 0; k < a.m; k++) {
    if (!cat[k]) return fail(CFB_ERR_INVALID, "categorical column %d is NULL", k);
    a.cols.cat[k] = cat[k];
    aligned = aligned && ((uintptr_t)cat[k] & 15) == 0;
  }
  a.cols.group = mask;
  a.n_rows = rows;
  a.mode = mode;
  a.out = d_out;
  const int device = M->device;
  const int smem_max = dev_info(device).smem_optin - 1024;
  const bool single = a.n_out == 1;
  if (single && mode == CFB_PREDICT_ARGMAX) {  // one output: the index of the largest score is 0
    if (mask) return fail(CFB_ERR_INVALID, "argmax of a single-output model with a row mask is not supported");
    CU(cudaMemsetAsync(d_out, 0, rows * 4, s));
    return CFB_OK;
  }
  a.vec4 = single && aligned;
  const size_t units = a.vec4 ? std::max<size_t>(1, rows / 4) : rows;
  int per_sm = (int)std::max<size_t>(1, std::min<size_t>(single ? 4 : 2, (size_t)smem_max / std::max<size_t>(M->smem, 1)));
  // a model that leaves room for at most two CTAs per SM: one CTA of 1024 threads shares it among 32 warps
  int threads = cfb::kPredictThreads;
  if (!single && cfb::predict_multi_max_threads(a.kb) >= 1024 && (size_t)dev_info(device).smem_sm / std::max<size_t>(M->smem + 1024, 1) <= 3 &&
      !getenv("CFB_PREDICT_SMALL_CTAS")) {
    threads = 1024;
    per_sm = 1;
  }
  const int grid = (int)std::max<size_t>(1, std::min<size_t>((size_t)dev_info(device).sms * per_sm, (units + threads - 1) / threads));
  auto run = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
    if (e != cudaSuccess) return e;
    kern<<<grid, threads, M->smem, s>>>(a);
    return cudaGetLastError();
  };
  cudaError_t e;
  switch (a.kb) {
    case 1: e = run(cfb::predict_score_kernel); break;
    case 2: e = run(cfb::predict_multi_kernel<2>); break;
    case 4: e = run(cfb::predict_multi_kernel<4>); break;
    case 6: e = run(cfb::predict_multi_kernel<6>); break;
    case 8: e = run(cfb::predict_multi_kernel<8>); break;
    case 10: e = run(cfb::predict_multi_kernel<10>); break;
    case 12: e = run(cfb::predict_multi_kernel<12>); break;
    case 16: e = run(cfb::predict_multi_kernel<16>); break;
    case 24: e = run(cfb::predict_multi_kernel<24>); break;
    default: e = run(cfb::predict_multi_kernel<32>); break;
  }
  if (e != cudaSuccess) return fail(CFB_ERR_CUDA, "predict kernel launch: %s", cudaGetErrorString(e));
  g_launches++;
  CU(cudaGetLastError());
  return CFB_OK;
}

// per host thread: a stream, a pinned tile and its device twin for cfb_predict_host
struct PredictScratch {
  int device = -1;
  cudaStream_t stream = nullptr;
  char *h = nullptr, *d = nullptr;
  size_t bytes = 0;
  ~PredictScratch() {
    if (device < 0 || device_count_quiet() == 0) return;
    cudaSetDevice(device);
    if (h) cudaFreeHost(h);
    if (d) cudaFree(d);
    if (stream) cudaStreamDestroy(stream);
  }
};
thread_local PredictScratch t_predict;

}  // namespace

extern "C" int cfb_model_create(int device, const cfb_linear_model *M, cfb_model **out_model) {
  if (!out_model) return fail(CFB_ERR_INVALID, "out is NULL");
  *out_model = nullptr;
  if (device_count_quiet() == 0) return fail(CFB_ERR_NO_DEVICE, "no CUDA device is visible; this library has no CPU fallback");
  if (!M) return fail(CFB_ERR_INVALID, "model is NULL");
  if (M->n_num < 0 || M->n_num > CFB_MAX_NUM || M->n_cat < 0 || M->n_cat > CFB_MAX_CAT)
    return fail(CFB_ERR_INVALID, "model: n_num / n_cat out of range");
  if (M->n_out < 1 || M->n_out > cfb::kPredictMaxOut) return fail(CFB_ERR_INVALID, "model: n_out must be in [1, %d]", cfb::kPredictMaxOut);
  if (!M->bias || (M->n_num && !M->w_num) || (M->n_cat && (!M->cat_offsets || !M->cat_keys || !M->w_cat)))
    return fail(CFB_ERR_INVALID, "model: NULL array");
  const int K = M->n_out, n = M->n_num, m = M->n_cat;
  const long long total = m ? M->cat_offsets[m] : 0;
  std::unique_ptr<cfb_model> h(new cfb_model);
  h->device = device;
  cfb::PredictArgs &a = h->args;
  a.n = n;
  a.m = m;
  a.n_out = K;
  a.total = (int)total;
  std::vector<int> map;  // dense key -> position maps
  for (int c = 0; c < m; c++) {
    const long long b = M->cat_offsets[c], e = M->cat_offsets[c + 1];
    if (e < b) return fail(CFB_ERR_INVALID, "model: cat_offsets must be non-decreasing");
    a.col_off[c] = (int)b;
    a.map_off[c] = (int)map.size();
    a.map_lo[c] = 0;
    a.map_len[c] = 0;
    if (e == b) continue;
    for (long long t = b + 1; t < e; t++)
      if (M->cat_keys[t] <= M->cat_keys[t - 1]) return fail(CFB_ERR_INVALID, "model: keys of column %d are not ascending", c);
    const long long len = (long long)M->cat_keys[e - 1] - M->cat_keys[b] + 1;
    if (len > (1 << 20) || (long long)map.size() + len > (4 << 20))
      return fail(CFB_ERR_DOMAIN, "model: the key range of column %d is too wide for the dense key map", c);
    a.map_lo[c] = M->cat_keys[b];
    a.map_len[c] = (int)len;
    map.resize(map.size() + len, -1);
    for (long long t = b; t < e; t++) map[a.map_off[c] + (M->cat_keys[t] - M->cat_keys[b])] = (int)t;
  }
  a.map_total = (int)map.size();
  // outputs padded to a kernel bucket; weights class-minor: [kb] bias | [n][kb] | [total][kb]
  int KB = 1;
  if (K > 1)
    for (int b : {2, 4, 6, 8, 10, 12, 16, 24, 32})
      if (b >= K) {
        KB = b;
        break;
      }
  a.kb = KB;
  h->smem = cfb::predict_smem_bytes(n, KB, (int)total, a.map_total);
  if (h->smem > (size_t)dev_info(device).smem_optin - 1024)
    return fail(CFB_ERR_DOMAIN, "model too large for the device path (%zu bytes of shared memory)", h->smem);
  std::vector<double> model((size_t)KB * (1 + n + total), 0.0);
  for (int k = 0; k < KB; k++) model[k] = k < K ? M->bias[k] : -std::numeric_limits<double>::infinity();
  for (int k = 0; k < K; k++) {
    for (int i = 0; i < n; i++) model[(size_t)KB * (1 + i) + k] = M->w_num[(size_t)k * n + i];
    for (long long t = 0; t < total; t++) model[(size_t)KB * (1 + n + t) + k] = M->w_cat[(size_t)k * total + t];
  }
  CU(cudaSetDevice(device));
  CU(cudaMalloc((void **)&h->d_model, model.size() * 8));
  if (cudaMalloc((void **)&h->d_map, std::max<size_t>(1, map.size()) * 4) != cudaSuccess) {
    cudaFree(h->d_model);
    return fail(CFB_ERR_CUDA, "cudaMalloc of the key map failed");
  }
  cudaError_t e = cudaMemcpy(h->d_model, model.data(), model.size() * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && !map.empty()) e = cudaMemcpy(h->d_map, map.data(), map.size() * 4, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(h->d_model);
    cudaFree(h->d_map);
    return fail(CFB_ERR_CUDA, "model upload: %s", cudaGetErrorString(e));
  }
  a.d_model = h->d_model;
  a.d_map = h->d_map;
  *out_model = h.release();
  return CFB_OK;
}

extern "C" void cfb_model_destroy(cfb_model *M) {
  if (!M) return;
  if (device_count_quiet() > 0) {
    cudaSetDevice(M->device);
    cudaFree(M->d_model);
    cudaFree(M->d_map);
    cudaFree(M->d_labels);
  }
  delete M;
}

namespace {
// Dense key -> position maps of a model's categorical columns (shared by the three model kinds).
int build_key_maps(int m, const int64_t *offs, const int32_t *keys, int *map_lo, int *map_off, int *map_len, std::vector<int> *map) {
  for (int c = 0; c < m; c++) {
    const long long b = offs[c], e = offs[c + 1];
    if (e < b) return fail(CFB_ERR_INVALID, "model: cat_offsets must be non-decreasing");
    map_off[c] = (int)map->size();
    map_lo[c] = 0;
    map_len[c] = 0;
    if (e == b) continue;
    for (long long t = b + 1; t < e; t++)
      if (keys[t] <= keys[t - 1]) return fail(CFB_ERR_INVALID, "model: keys of column %d are not ascending", c);
    const long long len = (long long)keys[e - 1] - keys[b] + 1;
    if (len > (1 << 20) || (long long)map->size() + len > (4 << 20))
      return fail(CFB_ERR_DOMAIN, "model: the key range of column %d is too wide for the dense key map", c);
    map_lo[c] = keys[b];
    map_len[c] = (int)len;
    map->resize(map->size() + len, -1);
    for (long long t = b; t < e; t++) (*map)[map_off[c] + (keys[t] - keys[b])] = (int)t;
  }
  return CFB_OK;
}

// doubles | key map | labels of a naive-Bayes / QDA model -> device
int upload_label_model(cfb_model *h, const std::vector<double> &vals, const std::vector<int> &map, const int32_t *labels, int K) {
  CU(cudaSetDevice(h->device));
  CU(cudaMalloc((void **)&h->d_model, std::max<size_t>(1, vals.size()) * 8));
  CU(cudaMalloc((void **)&h->d_map, std::max<size_t>(1, map.size()) * 4));
  CU(cudaMalloc((void **)&h->d_labels, (size_t)K * 4));
  if (!vals.empty()) CU(cudaMemcpy(h->d_model, vals.data(), vals.size() * 8, cudaMemcpyHostToDevice));
  if (!map.empty()) CU(cudaMemcpy(h->d_map, map.data(), map.size() * 4, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->d_labels, labels, (size_t)K * 4, cudaMemcpyHostToDevice));
  return CFB_OK;
}
}  // namespace

extern "C" int cfb_model_set_noise(cfb_model *M, double sigma, uint64_t seed, uint64_t first_row) {
  if (!M) return fail(CFB_ERR_INVALID, "model is NULL");
  if (M->type != 0 || M->args.n_out != 1) return fail(CFB_ERR_INVALID, "noise applies to a single-output linear model");
  if (!(sigma >= 0.0)) return fail(CFB_ERR_INVALID, "sigma must be >= 0");
  M->args.noise_sigma = sigma;
  M->args.noise_seed = seed;
  M->args.noise_first = first_row;
  return CFB_OK;
}

extern "C" int cfb_model_create_nb(int device, const cfb_nb_model *M, cfb_model **out_model) {
  if (!out_model) return fail(CFB_ERR_INVALID, "out is NULL");
  *out_model = nullptr;
  if (device_count_quiet() == 0) return fail(CFB_ERR_NO_DEVICE, "no CUDA device is visible; this library has no CPU fallback");
  if (!M) return fail(CFB_ERR_INVALID, "model is NULL");
  const int K = M->n_classes, n = M->n_num, m = M->n_cat;
  if (n < 0 || n > CFB_MAX_NUM || m < 0 || m > CFB_MAX_CAT || K < 1) return fail(CFB_ERR_INVALID, "model: shape out of range");
  if (!M->labels || !M->prior || (n && (!M->mean || !M->var)) || (m && (!M->cat_offsets || !M->cat_keys || !M->cat_prob)))
    return fail(CFB_ERR_INVALID, "model: NULL array");
  const long long total = m ? M->cat_offsets[m] : 0;
  std::unique_ptr<cfb_model> h(new cfb_model);
  h->device = device;
  h->type = 1;
  cfb::NbArgs &a = h->nb;
  a.n = n;
  a.m = m;
  a.n_classes = K;
  a.total = (int)total;
  std::vector<int> map;
  int rc = build_key_maps(m, M->cat_offsets, M->cat_keys, a.map_lo, a.map_off, a.map_len, &map);
  if (rc) return rc;
  // [K] prior | [K][n] norm | [K][n] mean | [K][n] 2 var | [K][total] probabilities: the row-independent factors of the
  // reference's expression, evaluated once with the same operations (naive_bayes.cpp:222-227)
  std::vector<double> v((size_t)K * (1 + 3 * n + total));
  double *prior = v.data(), *norm = prior + K, *mean = norm + (size_t)K * n, *two_var = mean + (size_t)K * n, *prob = two_var + (size_t)K * n;
  for (int k = 0; k < K; k++) {
    prior[k] = M->prior[k];
    for (int j = 0; j < n; j++) {
      double variance = M->var[(size_t)k * n + j];
      variance += 0.000000001;  // avoid division by 0 (naive_bayes.cpp:223)
      norm[(size_t)k * n + j] = (double)1 / sqrt(2 * M_PI * variance);
      mean[(size_t)k * n + j] = M->mean[(size_t)k * n + j];
      two_var[(size_t)k * n + j] = (double)2 * variance;
    }
    for (long long t = 0; t < total; t++) prob[(size_t)k * total + t] = M->cat_prob[(size_t)k * total + t];
  }
  rc = upload_label_model(h.get(), v, map, M->labels, K);
  if (rc) {
    cfb_model_destroy(h.release());
    return rc;
  }
  a.d_map = h->d_map;
  a.labels = h->d_labels;
  a.prior = h->d_model;
  a.norm = a.prior + K;
  a.mean = a.norm + (size_t)K * n;
  a.two_var = a.mean + (size_t)K * n;
  a.cat_prob = a.two_var + (size_t)K * n;
  *out_model = h.release();
  return CFB_OK;
}

extern "C" int cfb_model_create_qda(int device, const cfb_qda_model *M, cfb_model **out_model) {
  if (!out_model) return fail(CFB_ERR_INVALID, "out is NULL");
  *out_model = nullptr;
  if (device_count_quiet() == 0) return fail(CFB_ERR_NO_DEVICE, "no CUDA device is visible; this library has no CPU fallback");
  if (!M) return fail(CFB_ERR_INVALID, "model is NULL");
  const int K = M->n_classes, n = M->n_num, m = M->n_cat;
  if (n < 0 || n > CFB_MAX_NUM || m < 0 || m > CFB_MAX_CAT || K < 1) return fail(CFB_ERR_INVALID, "model: shape out of range");
  if (!M->labels || !M->quad || !M->lin || !M->intercept || (m && (!M->cat_offsets || !M->cat_keys))) return fail(CFB_ERR_INVALID, "model: NULL array");
  const long long total = m ? M->cat_offsets[m] : 0, P = n + total;
  if (P < 1 || (double)K * P * P * 8 > 4e9) return fail(CFB_ERR_DOMAIN, "model: QDA matrices of %lld x %lld per class are too large", P, P);
  std::unique_ptr<cfb_model> h(new cfb_model);
  h->device = device;
  h->type = 2;
  cfb::QdaArgs &a = h->qda;
  a.n = n;
  a.m = m;
  a.n_classes = K;
  a.total = (int)total;
  a.p = (int)P;
  std::vector<int> map;
  int rc = build_key_maps(m, M->cat_offsets, M->cat_keys, a.map_lo, a.map_off, a.map_len, &map);
  if (rc) return rc;
  // [K][P][P] Q | [K][P] g | [K] b with the centre folded in: g = lin - (Q + Q^T) c, b = intercept + c^T Q c - lin . c
  std::vector<double> v((size_t)K * (P * P + P + 1));
  double *Q = v.data(), *g = Q + (size_t)K * P * P, *b = g + (size_t)K * P;
  memcpy(Q, M->quad, (size_t)K * P * P * 8);
  for (int k = 0; k < K; k++) {
    const double *Qk = M->quad + (size_t)k * P * P, *lk = M->lin + (size_t)k * P;
    double bk = M->intercept[k];
    for (long long i = 0; i < P; i++) {
      double gi = lk[i];
      if (M->center) {
        for (long long j = 0; j < P; j++) gi -= (Qk[i + j * P] + Qk[j + i * P]) * M->center[j];
        double row = 0.0;
        for (long long j = 0; j < P; j++) row += Qk[i + j * P] * M->center[j];
        bk += M->center[i] * row - lk[i] * M->center[i];
      }
      g[(size_t)k * P + i] = gi;
    }
    b[k] = bk;
  }
  rc = upload_label_model(h.get(), v, map, M->labels, K);
  if (rc) {
    cfb_model_destroy(h.release());
    return rc;
  }
  a.d_map = h->d_map;
  a.labels = h->d_labels;
  a.Q = h->d_model;
  a.g = a.Q + (size_t)K * P * P;
  a.b = a.g + (size_t)K * P;
  *out_model = h.release();
  return CFB_OK;
}

extern "C" int cfb_predict_device(cfb_model *M, const float *const *d_num_cols, const int32_t *const *d_cat_cols,
                                  const int32_t *d_row_mask, size_t n_rows, int mode, void *d_out, void *stream) {
  NvtxRange nvtx_range("cfb_predict_device");
  if (!M) return fail(CFB_ERR_INVALID, "model is NULL");
  if (n_rows == 0) return CFB_OK;
  const int mn = M->type == 0 ? M->args.n : (M->type == 1 ? M->nb.n : M->qda.n), mm = M->type == 0 ? M->args.m : (M->type == 1 ? M->nb.m : M->qda.m);
  if (!d_out || (mn && !d_num_cols) || (mm && !d_cat_cols)) return fail(CFB_ERR_INVALID, "NULL column array / output");
  CU(cudaSetDevice(M->device));
  return predict_launch(M, d_num_cols, d_cat_cols, d_row_mask, n_rows, mode, d_out, (cudaStream_t)stream);
}

extern "C" int cfb_predict_host(cfb_model *M, const float *const *num_cols, const uint32_t *const *num_sel,
                                const int32_t *const *cat_cols, const uint32_t *const *cat_sel, size_t count, int mode,
                                void *out) {
  NvtxRange nvtx_range("cfb_predict_host");
  if (!M) return fail(CFB_ERR_INVALID, "model is NULL");
  if (count == 0) return CFB_OK;
  const int n = M->type == 0 ? M->args.n : (M->type == 1 ? M->nb.n : M->qda.n), m = M->type == 0 ? M->args.m : (M->type == 1 ? M->nb.m : M->qda.m);
  const int device = M->device;
  if (!out || (n && !num_cols) || (m && !cat_cols)) return fail(CFB_ERR_INVALID, "NULL argument");
  CU(cudaSetDevice(device));
  PredictScratch &sc = t_predict;
  if (sc.device != device) {
    if (sc.device >= 0) return fail(CFB_ERR_INVALID, "cfb_predict_host: a host thread stays on one device");
    sc.device = device;
    CU(cudaStreamCreateWithFlags(&sc.stream, cudaStreamNonBlocking));
  }
  const size_t rows = (count + 3) & ~(size_t)3, need = (size_t)(n + m + 1) * rows * 4;
  if (need > sc.bytes) {
    CU(cudaStreamSynchronize(sc.stream));
    if (sc.h) cudaFreeHost(sc.h);
    if (sc.d) cudaFree(sc.d);
    sc.h = sc.d = nullptr;
    sc.bytes = std::max<size_t>(need, (size_t)1 << 20);
    CU(cudaMallocHost((void **)&sc.h, sc.bytes));
    CU(cudaMalloc((void **)&sc.d, sc.bytes));
  }
  const float *dn[CFB_MAX_NUM];
  const int32_t *dc[CFB_MAX_CAT];
  for (int k = 0; k < n; k++) {
    gather((float *)(sc.h + (size_t)k * rows * 4), num_cols[k], num_sel ? num_sel[k] : nullptr, 0, count);
    dn[k] = (const float *)(sc.d + (size_t)k * rows * 4);
  }
  for (int k = 0; k < m; k++) {
    gather((int32_t *)(sc.h + (size_t)(n + k) * rows * 4), cat_cols[k], cat_sel ? cat_sel[k] : nullptr, 0, count);
    dc[k] = (const int32_t *)(sc.d + (size_t)(n + k) * rows * 4);
  }
  _mm_sfence();  // the tile was filled with non-temporal stores
  if (n + m) CU(cudaMemcpyAsync(sc.d, sc.h, (size_t)(n + m) * rows * 4, cudaMemcpyHostToDevice, sc.stream));
  char *d_out = sc.d + (size_t)(n + m) * rows * 4, *h_out = sc.h + (size_t)(n + m) * rows * 4;
  int rc = predict_launch(M, dn, dc, nullptr, count, mode, d_out, sc.stream);
  if (rc) return rc;
  CU(cudaMemcpyAsync(h_out, d_out, count * 4, cudaMemcpyDeviceToHost, sc.stream));
  CU(cudaStreamSynchronize(sc.stream));
  memcpy(out, h_out, count * 4);
  return CFB_OK;
}

// ======================================================================== C ABI
extern "C" {

int cfb_abi_version(void) { return CFB_ABI_VERSION; }
int cfb_device_count(void) { return device_count_quiet(); }
const char *cfb_last_error(void) { return g_err.c_str(); }
uint64_t cfb_kernel_launches(void) { return g_launches.load(); }
int cfb_set_timing(int enabled) {
  g_timing = enabled ? 1 : 0;
  return CFB_OK;
}

int cfb_ctx_create(int device, int kind, int n_num, int n_cat, int n_groups, cfb_ctx **out) {
  if (!out) return fail(CFB_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (kind != CFB_TRIPLE && kind != CFB_NB) return fail(CFB_ERR_INVALID, "unknown kind %d", kind);
  if (n_num < 0 || n_num > CFB_MAX_NUM || n_cat < 0 || n_cat > CFB_MAX_CAT || n_groups < 1)
    return fail(CFB_ERR_INVALID, "bad shape n_num=%d n_cat=%d n_groups=%d", n_num, n_cat, n_groups);
  const int nd = device_count_quiet();
  if (nd == 0) return fail(CFB_ERR_NO_DEVICE, "no CUDA device is visible; this library has no CPU fallback");
  if (device < 0 || device >= nd) return fail(CFB_ERR_INVALID, "device %d out of range (0..%d)", device, nd - 1);
  CU(cudaSetDevice(device));
  if (cfb_ctx *r = g_ctx_pool.take(device, kind, n_num, n_cat, n_groups)) {
    r->timed = g_timing.load() != 0;
    *out = r;
    return CFB_OK;
  }
  cfb_ctx *c = new cfb_ctx();
  c->device = device;
  c->kind = kind;
  c->n = n_num;
  c->m = n_cat;
  c->G = n_groups;
  c->timed = g_timing.load() != 0;
  auto bail = [&](int rc) {
    cfb_ctx_destroy(c);
    return rc;
  };
#define CUB(call)                                                                              \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
      return bail(fail(e_ == cudaErrorMemoryAllocation ? CFB_ERR_OOM : CFB_ERR_CUDA, "%s: %s", #call, \
                       cudaGetErrorString(e_)));                                               \
  } while (0)
  CUB(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  build_layout(c->lay, kind, n_num, n_cat, n_groups, nullptr, nullptr);
  int rc = alloc_state(c->lay, &c->d_f64, &c->d_u64, c->stream);
  if (rc) return bail(rc);
  CUB(cudaMalloc(&c->d_lay, sizeof(Layout)));
  CUB(cudaMemcpyAsync(c->d_lay, &c->lay, sizeof(Layout), cudaMemcpyHostToDevice, c->stream));
  CUB(cudaMalloc(&c->d_err, sizeof(int)));
  CUB(cudaMemsetAsync(c->d_err, 0, sizeof(int), c->stream));
  c->gram_grid = dev_info(device).sms;
  if (const char *e = getenv("CFB_GRAM_GRID")) c->gram_grid = std::max(1, atoi(e));
  CUB(cudaMalloc(&c->d_partials, (size_t)c->gram_grid * (n_num + n_num * (n_num + 1) / 2 + 1) * sizeof(double)));
  CUB(cudaMalloc(&c->d_ticket, sizeof(unsigned int)));
  CUB(cudaMemsetAsync(c->d_ticket, 0, sizeof(unsigned int), c->stream));
  CUB(cudaEventCreate(&c->ev0));
  CUB(cudaEventCreate(&c->ev1));
  CUB(cudaStreamSynchronize(c->stream));
#undef CUB
  *out = c;
  return CFB_OK;
}

int cfb_ctx_destroy(cfb_ctx *c) {
  if (!c) return CFB_OK;
  drop_reduced(c);
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->user_stream) cudaStreamSynchronize(c->user_stream);
  for (auto &s : c->st) g_stage_pool.release(c->device, s);  // the stream is drained: nothing in flight
  c->tile_rows = 0;
  c->fill = 0;
  c->cur = 0;
  c->uses_group = false;
  c->user_stream = nullptr;
  if (c->packed_cap > (160u << 20)) {  // a parked context keeps moderate scratch only (<= 13 M rows x 12 columns)
    g_packed_pool.give(c->device, c->d_packed, c->packed_cap);  // (the streams are drained: nothing reads it)
    c->d_packed = nullptr;
    c->packed_cap = 0;
  }
  // recycle: zero the state (its layout, i.e. the categorical domain seen so far, is kept: keys
  // that do not occur again have count 0 and are not emitted) and park the context
  const long long state_bytes = (c->lay.F + c->lay.U) * c->lay.n_groups * 8;
  if (c->stream && c->d_f64 && state_bytes <= CtxPool::kMaxStateBytes && !c->lay.pairs_hashed && !any_dict(c) &&
      !getenv("CFB_NO_CTX_POOL")) {
    bool ok = cudaMemsetAsync(c->d_f64, 0, std::max<long long>(8, c->lay.F * c->lay.n_groups * 8), c->stream) == cudaSuccess &&
              cudaMemsetAsync(c->d_u64, 0, std::max<long long>(8, c->lay.U * c->lay.n_groups * 8), c->stream) == cudaSuccess &&
              cudaMemsetAsync(c->d_err, 0, sizeof(int), c->stream) == cudaSuccess &&
              cudaMemsetAsync(c->d_ticket, 0, sizeof(unsigned int), c->stream) == cudaSuccess &&
              cudaStreamSynchronize(c->stream) == cudaSuccess;
    c->user_domain = false;
    c->touched = false;
    if (ok && g_ctx_pool.give(c)) return CFB_OK;
    cudaGetLastError();
  }
  cudaFree(c->d_f64);
  cudaFree(c->d_u64);
  cudaFree(c->d_lay);
  cudaFree(c->d_err);
  cudaFree(c->d_partials);
  cudaFree(c->d_ticket);
  cudaFree(c->d_slab);
  cudaFree(c->d_chain_slab);
  cudaFree(c->d_cnt_slab);
  cudaFree(c->d_packed);
  delete c->role_plan;
  if (c->hash.capacity) hash_free(c->hash, c->G);
  for (auto &cd : c->dict)
    if (cd.on || cd.d.table) dict_free(cd);
  cudaFree(c->d_remap);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->stream) cudaStreamDestroy(c->stream);
  cudaGetLastError();
  delete c;
  return CFB_OK;
}

int cfb_ctx_set_cat_domain(cfb_ctx *c, const int32_t *lo, const int32_t *hi) {
  if (!c || !lo || !hi) return fail(CFB_ERR_INVALID, "NULL argument");
  if (int rc = check_open(c)) return rc;
  for (int k = 0; k < c->m; k++)
    if (hi[k] < lo[k]) return fail(CFB_ERR_INVALID, "empty domain for categorical column %d", k);
  CU(cudaSetDevice(c->device));
  int rc = ensure_domain(c, lo, hi, /*exact=*/!c->touched);
  if (rc) return rc;
  c->user_domain = true;
  return CFB_OK;
}

int cfb_ctx_append(cfb_ctx *c, const float *const *num_cols, const uint32_t *const *num_sel,
                   const int32_t *const *cat_cols, const uint32_t *const *cat_sel, const uint32_t *group_slot,
                   size_t count) {
  NvtxRange nvtx_range("cfb_ctx_append");
  if (!c) return fail(CFB_ERR_INVALID, "ctx is NULL");
  if (int rc = check_open(c)) return rc;
  if (count == 0) return CFB_OK;
  if ((c->n && !num_cols) || (c->m && !cat_cols)) return fail(CFB_ERR_INVALID, "column array is NULL");
  CU(cudaSetDevice(c->device));
  c->touched = true;
  int rc = ensure_staging(c);
  if (rc) return rc;
  if (group_slot && !c->uses_group) {
    if (c->fill) {  // rows staged so far had no slot column: they belong to slot 0
      memset(c->st[c->cur].h + (size_t)(c->n + c->m) * c->tile_rows * 4, 0, c->fill * 4);
    }
    c->uses_group = true;
  }
  size_t done = 0;
  while (done < count) {
    const size_t take = std::min(count - done, c->tile_rows - c->fill);
    char *base = c->st[c->cur].h;
    const size_t tr = c->tile_rows;
    for (int k = 0; k < c->n; k++)
      gather((float *)(base + (size_t)k * tr * 4) + c->fill, num_cols[k], num_sel ? num_sel[k] : nullptr, done, take);
    for (int k = 0; k < c->m; k++) {
      int32_t *dst = (int32_t *)(base + (size_t)(c->n + k) * tr * 4) + c->fill;
      const uint32_t *ks = cat_sel ? cat_sel[k] : nullptr;
      gather(dst, cat_cols[k], ks, done, take);
      if (!c->user_domain) {
        // key range of the tile: a contiguous vector is read from the source (in cache after the copy -- the tile itself
        // was written around the cache), a gathered one from the tile (written with plain stores)
        int lo = c->st_lo[k], hi = c->st_hi[k];
        cfb_stage_minmax32(ks ? dst : cat_cols[k] + done, take, &lo, &hi);
        c->st_lo[k] = lo;
        c->st_hi[k] = hi;
      }
    }
    if (c->uses_group) {
      uint32_t *dst = (uint32_t *)(base + (size_t)(c->n + c->m) * tr * 4) + c->fill;
      if (group_slot)
        memcpy(dst, group_slot + done, take * 4);
      else
        memset(dst, 0, take * 4);
    }
    c->fill += take;
    done += take;
    if (c->fill == c->tile_rows) {
      rc = flush_tile(c);
      if (rc) return rc;
    }
  }
  return CFB_OK;
}

int cfb_ctx_append_triples(cfb_ctx *c, size_t count, const int32_t *N, const float *lin, const float *quad,
                           const cfb_list_entry *lin_cat_lists, const int32_t *lc_key, const float *lc_val,
                           const cfb_list_entry *num_cat_lists, const int32_t *nc_key, const float *nc_val,
                           const cfb_list_entry *cat_cat_lists, const int32_t *cc_key1, const int32_t *cc_key2,
                           const float *cc_val) {
  return cfb_ctx_append_triples_slot(c, 0, count, N, lin, quad, lin_cat_lists, lc_key, lc_val, num_cat_lists, nc_key, nc_val,
                                     cat_cat_lists, cc_key1, cc_key2, cc_val);
}

int cfb_ctx_append_triples_slot(cfb_ctx *c, int slot, size_t count, const int32_t *N, const float *lin, const float *quad,
                                const cfb_list_entry *lin_cat_lists, const int32_t *lc_key, const float *lc_val,
                                const cfb_list_entry *num_cat_lists, const int32_t *nc_key, const float *nc_val,
                                const cfb_list_entry *cat_cat_lists, const int32_t *cc_key1, const int32_t *cc_key2,
                                const float *cc_val) {
  NvtxRange nvtx_range("cfb_ctx_append_triples_slot");
  if (!c) return fail(CFB_ERR_INVALID, "ctx is NULL");
  if (int rc = check_open(c)) return rc;
  if (slot < 0 || slot >= c->G) return fail(CFB_ERR_INVALID, "slot %d out of range", slot);
  if (count == 0) return CFB_OK;
  const int n = c->n, m = c->m;
  const size_t nq = (size_t)c->lay.nq;
  if (!N || (n && (!lin || !quad)) || (m && !lin_cat_lists)) return fail(CFB_ERR_INVALID, "NULL child array");
  if (c->kind == CFB_TRIPLE && m && (!cat_cat_lists || (n && !num_cat_lists)))
    return fail(CFB_ERR_INVALID, "NULL categorical child array");
  CU(cudaSetDevice(c->device));
  c->touched = true;
  int rc = flush_tile(c);  // keep the order of updates on the context stream
  if (rc) return rc;
  // sparse entries -> tagged records (host: a few entries per lifted row), key ranges for the domain
  std::vector<cfb::LiftedEntry> ent;
  int lo[cfb::kMaxCat], hi[cfb::kMaxCat];
  for (int k = 0; k < m; k++) {
    lo[k] = INT_MAX;
    hi[k] = INT_MIN;
  }
  auto see = [&](int col, int key) {
    lo[col] = std::min(lo[col], key);
    hi[col] = std::max(hi[col], key);
  };
  for (size_t r = 0; r < count; r++) {
    for (int k = 0; k < m; k++) {
      const cfb_list_entry le = lin_cat_lists[r * m + k];
      for (uint64_t t = le.offset; t < le.offset + le.length; t++) {
        ent.push_back({k, lc_key[t], 0, lc_val[t]});
        see(k, lc_key[t]);
      }
    }
    if (c->kind != CFB_TRIPLE) continue;
    for (int l = 0; l < n; l++)
      for (int k = 0; k < m; k++) {
        const cfb_list_entry le = num_cat_lists[(r * n + l) * m + k];
        for (uint64_t t = le.offset; t < le.offset + le.length; t++) {
          ent.push_back({64 + l * 32 + k, nc_key[t], 0, nc_val[t]});
          see(k, nc_key[t]);
        }
      }
    size_t p = 0;
    for (int k = 0; k < m; k++)
      for (int l = k; l < m; l++, p++) {
        if (k == l) continue;  // diagonal pair lists repeat lin_cat; the state derives them
        const cfb_list_entry le = cat_cat_lists[r * ((size_t)m * (m + 1) / 2) + p];
        for (uint64_t t = le.offset; t < le.offset + le.length; t++) {
          ent.push_back({2048 + k * 32 + l, cc_key1[t], cc_key2[t], cc_val[t]});
          see(k, cc_key1[t]);
          see(l, cc_key2[t]);
        }
      }
  }
  if (m > 0 && !c->user_domain && !ent.empty()) {
    for (int k = 0; k < m; k++)
      if (lo[k] > hi[k]) lo[k] = hi[k] = c->lay.has_domain ? c->lay.lo[k] : 0;
    // wide-range columns: switch to dictionary codes and rewrite this chunk's keys
    for (int k = 0; k < m; k++) {
      if (!c->dict[k].on) {
        long long l = lo[k], h = hi[k];
        if (c->lay.has_domain) {
          l = std::min<long long>(l, c->lay.lo[k]);
          h = std::max<long long>(h, (long long)c->lay.lo[k] + c->lay.dom[k] - 1);
        }
        if (h - l + 1 > dict_range_limit()) {
          rc = dict_enable(c, k, c->stream);
          if (rc) return rc;
        }
      }
      if (!c->dict[k].on) continue;
      std::vector<int32_t> keys;
      std::vector<std::pair<size_t, int>> where;  // (entry, 1 = key1 / 2 = key2)
      for (size_t i = 0; i < ent.size(); i++) {
        const int tag = ent[i].tag;
        if (tag < 64 ? tag == k : (tag < 2048 ? (tag - 64) % 32 == k : (tag - 2048) / 32 == k)) {
          keys.push_back(ent[i].key1);
          where.emplace_back(i, 1);
        }
        if (tag >= 2048 && (tag - 2048) % 32 == k) {
          keys.push_back(ent[i].key2);
          where.emplace_back(i, 2);
        }
      }
      rc = dict_codes_for_host_keys(c, k, keys);
      if (rc) return rc;
      for (size_t j = 0; j < where.size(); j++)
        (where[j].second == 1 ? ent[where[j].first].key1 : ent[where[j].first].key2) = keys[j];
      lo[k] = 0;
      hi[k] = std::max(c->dict[k].n_codes - 1, 0);
    }
    rc = ensure_domain(c, lo, hi);
    if (rc) return rc;
  }
  // device scratch: N | lin | quad | entries
  const size_t bN = count * 4, bL = count * n * 4, bQ = count * nq * 4, bE = ent.size() * sizeof(cfb::LiftedEntry);
  auto up16 = [](size_t b) { return (b + 15) & ~(size_t)15; };
  char *d = nullptr;
  CU(cudaMalloc((void **)&d, std::max<size_t>(16, up16(bN) + up16(bL) + up16(bQ) + up16(bE))));
  char *dN = d, *dL = dN + up16(bN), *dQ = dL + up16(bL), *dE = dQ + up16(bQ);
  cudaStream_t s = c->stream;
  CU(cudaMemcpyAsync(dN, N, bN, cudaMemcpyHostToDevice, s));
  if (bL) CU(cudaMemcpyAsync(dL, lin, bL, cudaMemcpyHostToDevice, s));
  if (bQ) CU(cudaMemcpyAsync(dQ, quad, bQ, cudaMemcpyHostToDevice, s));
  if (bE) CU(cudaMemcpyAsync(dE, ent.data(), bE, cudaMemcpyHostToDevice, s));
  // the state of GROUP BY slot `slot` (ensure_domain above may have re-laid out the state: read the layout now)
  double *sf64 = c->d_f64 + (long long)slot * c->lay.F;
  unsigned long long *su64 = c->d_u64 + (long long)slot * c->lay.U;
  cfb::lifted_count_kernel<<<1, 256, 0, s>>>((const int32_t *)dN, count, su64);
  g_launches++;
  if (n) {
    cfb::lifted_colsum_kernel<<<std::min(n, 64), 256, 0, s>>>((const float *)dL, count, n, sf64);
    cfb::lifted_colsum_kernel<<<(int)std::min<size_t>(nq, 128), 256, 0, s>>>((const float *)dQ, count, (int)nq, sf64 + n);
    g_launches += 2;
  }
  if (bE) {
    const int blocks = (int)std::min<size_t>((ent.size() + 255) / 256, (size_t)dev_info(c->device).sms * 4);
    rc = hash_reserve(c, ent.size());
    if (rc) return rc;
    cfb::lifted_scatter_kernel<<<blocks, 256, 0, s>>>((const cfb::LiftedEntry *)dE, ent.size(), c->d_lay, sf64, su64, c->d_err,
                                                     c->hash, slot);
    g_launches++;
  }
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s));  // the host arrays (DuckDB vectors, `ent`) go away after this call
  cudaFree(d);
  return CFB_OK;
}

int cfb_triple_device(cfb_ctx *c, const float *const *d_num_cols, const int32_t *const *d_cat_cols,
                      const int32_t *d_group_slot, size_t n_rows, void *stream) {
  NvtxRange nvtx_range("cfb_triple_device");
  if (!c) return fail(CFB_ERR_INVALID, "ctx is NULL");
  if (int rc = check_open(c)) return rc;
  if ((c->n && !d_num_cols) || (c->m && !d_cat_cols)) return fail(CFB_ERR_INVALID, "column array is NULL");
  CU(cudaSetDevice(c->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
  if (stream) c->user_stream = s;
  if (c->m == 0 || c->user_domain || n_rows == 0) return scan_device(c, d_num_cols, d_cat_cols, d_group_slot, n_rows, s);
  // No declared domain: one extra pass per slice finds [min,max] of every categorical column (and
  // feeds the key dictionaries of wide-range columns); slices bound the dictionary scratch.
  const size_t slice = (size_t)1 << 26;
  for (size_t r0 = 0; r0 < n_rows; r0 += slice) {
    const size_t cnt = std::min(slice, n_rows - r0);
    const float *num[CFB_MAX_NUM];
    const int32_t *cat[cfb::kMaxCat], *eff[cfb::kMaxCat];
    for (int k = 0; k < c->n; k++) num[k] = d_num_cols[k] + r0;
    for (int k = 0; k < c->m; k++) cat[k] = d_cat_cols[k] + r0;
    int lo[cfb::kMaxCat], hi[cfb::kMaxCat];
    int rc = cfb_cat_minmax_device(c->device, cat, c->m, cnt, lo, hi, s);
    if (rc) return rc;
    rc = prepare_cats(c, cat, cnt, s, lo, hi, /*in_place=*/false, eff);
    if (rc) return rc;
    rc = scan_device(c, num, eff, d_group_slot ? d_group_slot + r0 : nullptr, cnt, s);
    if (rc) return rc;
  }
  return CFB_OK;
}

int cfb_ctx_sync(cfb_ctx *c) {
  NvtxRange nvtx_range("cfb_ctx_sync");
  if (!c) return fail(CFB_ERR_INVALID, "ctx is NULL");
  CU(cudaSetDevice(c->device));
  int rc = flush_tile(c);
  if (rc) return rc;
  if (c->user_stream) CU(cudaStreamSynchronize(c->user_stream));
  return check_async_error(c);
}

double cfb_last_scan_ms(cfb_ctx *c) {
  if (!c || !c->timed) return -1.0;
  float ms = 0.f;
  if (cudaEventSynchronize(c->ev1) != cudaSuccess || cudaEventElapsedTime(&ms, c->ev0, c->ev1) != cudaSuccess) {
    cudaGetLastError();
    return -1.0;
  }
  return (double)ms;
}

namespace {
// dst[dst_slots[i]] += src[src_slots[i]] inside ONE context (two GROUP BY states of the same arena: DuckDB's
// radix-partitioned hash aggregate can emit a group twice from one thread and combine the two).  A slot may
// not be source and target in the same call: the adds run concurrently.
int combine_within(cfb_ctx *c, size_t n_pairs, const int32_t *dst_slots, const int32_t *src_slots) {
  if (!n_pairs) return fail(CFB_ERR_INVALID, "cannot combine a context with itself (name the slots)");
  if (!dst_slots || !src_slots) return fail(CFB_ERR_INVALID, "slot arrays are NULL");
  std::vector<int> gmap(c->G, -1);
  std::vector<char> is_dst(c->G, 0);
  for (size_t i = 0; i < n_pairs; i++) {
    const int d = dst_slots[i], s = src_slots[i];
    if (s < 0 || s >= c->G || d < 0 || d >= c->G) return fail(CFB_ERR_INVALID, "combine: slot out of range");
    if (s == d) return fail(CFB_ERR_INVALID, "combine: a slot cannot be merged into itself");
    if (gmap[s] != -1 || is_dst[d]) return fail(CFB_ERR_INVALID, "combine: a slot appears twice");
    gmap[s] = d;
    is_dst[d] = 1;
  }
  for (int g = 0; g < c->G; g++)
    if (is_dst[g] && gmap[g] != -1) return fail(CFB_ERR_INVALID, "combine: a slot is both source and target");
  int rc = cfb_ctx_sync(c);
  if (rc) return rc;
  CU(cudaSetDevice(c->device));
  int *d_gmap = nullptr;
  CU(cudaMalloc(&d_gmap, gmap.size() * sizeof(int)));
  CU(cudaMemcpyAsync(d_gmap, gmap.data(), gmap.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  rc = launch_remap_add(c->d_lay, c->d_lay, c->lay, c->d_f64, c->d_u64, c->d_f64, c->d_u64, c->hash, c->d_err, c->stream,
                        cfb::SlotTrans{}, d_gmap);
  if (rc == CFB_OK && c->lay.pairs_hashed) {
    // every pair of a source partition may be new to its target partition
    unsigned long long exact = 0;
    CU(cudaMemcpyAsync(&exact, c->hash.n_entries, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    rc = hash_reserve(c, exact);
    if (rc == CFB_OK) {
      cfb::pair_hash_drain_kernel<<<148 * 4, 256, 0, c->stream>>>(c->hash, c->d_lay, c->d_lay, c->d_u64, c->hash, c->d_err,
                                                                 cfb::SlotTrans{}, d_gmap, c->G);
      g_launches++;
      if (cudaGetLastError() != cudaSuccess) rc = fail(CFB_ERR_CUDA, "pair hash drain launch failed");
    }
  }
  cudaStreamSynchronize(c->stream);
  cudaFree(d_gmap);
  return rc;
}
}  // namespace

int cfb_ctx_combine(cfb_ctx *dst, const cfb_ctx *src) {
  if (dst && src && dst->G != src->G) return fail(CFB_ERR_INVALID, "combine: group counts differ (use cfb_ctx_combine_slots)");
  return cfb_ctx_combine_slots(dst, src, 0, nullptr, nullptr);
}

int cfb_ctx_combine_slots(cfb_ctx *dst, const cfb_ctx *src_c, size_t n_pairs, const int32_t *dst_slots,
                          const int32_t *src_slots) {
  NvtxRange nvtx_range("cfb_ctx_combine_slots");
  cfb_ctx *src = const_cast<cfb_ctx *>(src_c);
  if (!dst || !src) return fail(CFB_ERR_INVALID, "ctx is NULL");
  if (int rc = check_open(dst)) return rc;
  if (int rc = check_open(src)) return rc;
  if (dst == src) return combine_within(dst, n_pairs, dst_slots, src_slots);
  if (dst->kind != src->kind || dst->n != src->n || dst->m != src->m) return fail(CFB_ERR_INVALID, "combine: shapes differ");
  // slot map: dst slot of every src slot (-1 = not combined); identity when no pairs are given
  std::vector<int> gmap;
  if (n_pairs) {
    if (!dst_slots || !src_slots) return fail(CFB_ERR_INVALID, "slot arrays are NULL");
    gmap.assign(src->G, -1);
    std::vector<char> used(dst->G, 0);
    for (size_t i = 0; i < n_pairs; i++) {
      if (src_slots[i] < 0 || src_slots[i] >= src->G || dst_slots[i] < 0 || dst_slots[i] >= dst->G)
        return fail(CFB_ERR_INVALID, "combine: slot out of range");
      if (gmap[src_slots[i]] != -1 || used[dst_slots[i]]) return fail(CFB_ERR_INVALID, "combine: a slot appears twice");
      gmap[src_slots[i]] = dst_slots[i];
      used[dst_slots[i]] = 1;
    }
  }
  int rc = cfb_ctx_sync(src);
  if (rc) return rc;
  CU(cudaSetDevice(dst->device));
  dst->touched = true;
  rc = flush_tile(dst);
  if (rc) return rc;
  const double *sf = src->d_f64;
  const unsigned long long *su = src->d_u64;
  const Layout *sl = src->d_lay;
  double *tf = nullptr;
  unsigned long long *tu = nullptr;
  Layout *tl = nullptr;
  if (src->device != dst->device) {
    // stage the source state on dst's device (peer copy; the driver routes it over NVLink)
    const size_t bf = std::max<long long>(8, src->lay.F * src->lay.n_groups * 8);
    const size_t bu = std::max<long long>(8, src->lay.U * src->lay.n_groups * 8);
    CU(cudaMalloc(&tf, bf));
    CU(cudaMalloc(&tu, bu));
    CU(cudaMalloc(&tl, sizeof(Layout)));
    CU(cudaMemcpyPeerAsync(tf, dst->device, src->d_f64, src->device, bf, dst->stream));
    CU(cudaMemcpyPeerAsync(tu, dst->device, src->d_u64, src->device, bu, dst->stream));
    CU(cudaMemcpyAsync(tl, &src->lay, sizeof(Layout), cudaMemcpyHostToDevice, dst->stream));
    sf = tf;
    su = tu;
    sl = tl;
  }
  cfb::SlotTrans tr{};
  std::vector<void *> scratch;  // translation arrays and staged dictionaries, freed at the end
  if (src->m > 0 && src->lay.has_domain) {
    int lo[cfb::kMaxCat], hi[cfb::kMaxCat];
    for (int k = 0; k < src->m; k++) {
      lo[k] = src->lay.lo[k];
      hi[k] = (int)((long long)src->lay.lo[k] + src->lay.dom[k] - 1);
      // a column keyed by a dictionary on either side (or too wide once united) is merged by KEY:
      // dst gets a dictionary, and the source's slots are translated to dst codes
      bool by_key = src->dict[k].on || dst->dict[k].on;
      if (!by_key && dst->lay.has_domain && !dst->user_domain) {
        const long long l = std::min<long long>(lo[k], dst->lay.lo[k]);
        const long long h = std::max<long long>(hi[k], (long long)dst->lay.lo[k] + dst->lay.dom[k] - 1);
        by_key = h - l + 1 > dict_range_limit();
      } else if (!by_key && !dst->user_domain) {
        by_key = (long long)hi[k] - lo[k] + 1 > dict_range_limit();
      }
      if (!by_key) continue;
      rc = dict_enable(dst, k, dst->stream);
      if (rc) return rc;
      const long long n_slots = src->dict[k].on ? src->dict[k].n_codes : src->lay.dom[k];
      if (n_slots == 0) {
        lo[k] = 0;
        hi[k] = std::max(dst->dict[k].n_codes - 1, 0);
        continue;
      }
      const int *keys_src = nullptr;
      if (src->dict[k].on) {
        keys_src = src->dict[k].d.keys_of_code;
        if (src->device != dst->device) {
          int *tmp = nullptr;
          CU(cudaMalloc(&tmp, n_slots * sizeof(int)));
          CU(cudaMemcpyPeerAsync(tmp, dst->device, keys_src, src->device, n_slots * sizeof(int), dst->stream));
          scratch.push_back(tmp);
          keys_src = tmp;
        }
      }
      rc = dict_reserve(dst, k, (unsigned long long)n_slots, dst->stream);
      if (rc) return rc;
      int *trans = nullptr;
      CU(cudaMalloc(&trans, n_slots * sizeof(int)));
      scratch.push_back(trans);
      const unsigned long long *counts = su + 1 + src->lay.cat_off[k];
      const int blocks = (int)std::min<long long>((n_slots + 255) / 256, 148 * 4);
      for (int phase = 0; phase < 2; phase++)
        cfb::dict_translate_kernel<<<blocks, 256, 0, dst->stream>>>(dst->dict[k].d, keys_src, src->lay.lo[k], counts, src->lay.U,
                                                                   src->G, n_slots, trans, phase);
      g_launches += 2;
      CU(cudaGetLastError());
      rc = dict_read_counts(dst, dst->stream);
      if (rc) return rc;
      tr.col[k] = trans;
      lo[k] = 0;
      hi[k] = std::max(dst->dict[k].n_codes - 1, 0);
    }
    rc = ensure_domain(dst, lo, hi);
    if (rc) return rc;
  }
  // pairs of the source that may be new to dst: its hash entries, or its dense non-zeros
  cfb::PairHash shash = src->hash, thash{};
  if (dst->lay.pairs_hashed) {
    unsigned long long add = (unsigned long long)dense_pair_entries(src->lay);
    if (src->lay.pairs_hashed) {
      unsigned long long exact = 0;
      CU(cudaSetDevice(src->device));
      CU(cudaMemcpy(&exact, src->hash.n_entries, 8, cudaMemcpyDeviceToHost));
      CU(cudaSetDevice(dst->device));
      add = exact;
    }
    rc = hash_reserve(dst, add);
    if (rc) return rc;
  }
  if (src->lay.pairs_hashed && src->device != dst->device) {
    const size_t bytes = (size_t)src->hash.capacity * src->G * 8;
    thash.capacity = src->hash.capacity;
    CU(cudaMalloc(&thash.keys, bytes));
    CU(cudaMalloc(&thash.counts, bytes));
    CU(cudaMemcpyPeerAsync(thash.keys, dst->device, src->hash.keys, src->device, bytes, dst->stream));
    CU(cudaMemcpyPeerAsync(thash.counts, dst->device, src->hash.counts, src->device, bytes, dst->stream));
    shash = thash;
  }
  int *d_gmap = nullptr;
  if (!gmap.empty()) {
    CU(cudaMalloc(&d_gmap, gmap.size() * sizeof(int)));
    CU(cudaMemcpyAsync(d_gmap, gmap.data(), gmap.size() * sizeof(int), cudaMemcpyHostToDevice, dst->stream));
    scratch.push_back(d_gmap);
  }
  rc = launch_remap_add(dst->d_lay, sl, src->lay, dst->d_f64, dst->d_u64, sf, su, dst->hash, dst->d_err, dst->stream, tr, d_gmap);
  if (rc) return rc;
  if (src->lay.pairs_hashed) {
    cfb::pair_hash_drain_kernel<<<148 * 4, 256, 0, dst->stream>>>(shash, dst->d_lay, sl, dst->d_u64, dst->hash, dst->d_err, tr,
                                                                   d_gmap, src->G);
    g_launches++;
    CU(cudaGetLastError());
  }
  CU(cudaStreamSynchronize(dst->stream));
  cudaFree(tf);
  cudaFree(tu);
  cudaFree(tl);
  cudaFree(thash.keys);
  cudaFree(thash.counts);
  for (void *p : scratch) cudaFree(p);
  return CFB_OK;
}

int cfb_ctx_finalize(cfb_ctx *c, int group, cfb_result *out) {
  NvtxRange nvtx_range("cfb_ctx_finalize");
  if (!c || !out) return fail(CFB_ERR_INVALID, "NULL argument");
  if (group < 0 || group >= c->G) return fail(CFB_ERR_INVALID, "group %d out of range", group);
  memset(out, 0, sizeof(*out));
  if (!c->reduced.empty()) return copy_result(&c->reduced[(size_t)group], out);
  int rc = cfb_ctx_sync(c);
  if (rc) return rc;
  // the aggregate is complete: hand the (drained) staging tiles back; a later append re-acquires them
  if (c->tile_rows && c->fill == 0) {
    for (auto &st : c->st) {
      if (st.in_flight) CU(cudaEventSynchronize(st.done));
      g_stage_pool.release(c->device, st);
    }
    c->tile_rows = 0;
    c->cur = 0;
  }
  const Layout &L = c->lay;
  std::vector<double> f(std::max<long long>(1, L.F));
  std::vector<unsigned long long> u(std::max<long long>(1, L.U));
  CU(cudaMemcpyAsync(f.data(), c->d_f64 + (long long)group * L.F, L.F * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(u.data(), c->d_u64 + (long long)group * L.U, L.U * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  const int n = c->n, m = c->m;
  out->kind = c->kind;
  out->n_num = n;
  out->n_cat = m;
  out->N = (int64_t)u[0];
  out->n_quad = L.nq;
  auto dupv = [](const auto &v) {
    using T = typename std::decay<decltype(v)>::type::value_type;
    T *p = (T *)malloc(std::max<size_t>(1, v.size()) * sizeof(T));
    if (!v.empty()) memcpy(p, v.data(), v.size() * sizeof(T));
    return p;
  };
  out->lin = dupv(std::vector<double>(f.begin(), f.begin() + n));
  out->quad = dupv(std::vector<double>(f.begin() + n, f.begin() + n + L.nq));
  // keys that occurred, ascending per column: the iteration order of the reference's std::map.
  // Dense columns: key = lo + slot (already ascending); dictionary columns: key = keys_of_code[slot],
  // sorted here.
  std::vector<int32_t> keys;
  std::vector<int64_t> counts, offs(m + 1, 0), dense_t;
  std::vector<std::vector<int32_t>> koc(m);  // keys of codes, dictionary columns only
  for (int k = 0; k < m; k++)
    if (c->dict[k].on && c->dict[k].n_codes > 0) {
      koc[k].resize(c->dict[k].n_codes);
      CU(cudaMemcpyAsync(koc[k].data(), c->dict[k].d.keys_of_code, koc[k].size() * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    }
  CU(cudaStreamSynchronize(c->stream));
  auto key_of = [&](int col, long long slot) -> int32_t {
    return c->dict[col].on ? koc[col][(size_t)slot] : (int32_t)((long long)L.lo[col] + slot);
  };
  for (int k = 0; k < m; k++) {
    std::vector<std::pair<int32_t, int>> occ;  // (key, slot)
    for (int s = 0; s < L.dom[k]; s++)
      if (u[1 + L.cat_off[k] + s] && (!c->dict[k].on || s < (int)koc[k].size())) occ.emplace_back(key_of(k, s), s);
    if (c->dict[k].on) std::sort(occ.begin(), occ.end());
    for (const auto &ks : occ) {
      keys.push_back(ks.first);
      counts.push_back((int64_t)u[1 + L.cat_off[k] + ks.second]);
      dense_t.push_back(L.cat_off[k] + ks.second);
    }
    offs[k + 1] = (int64_t)keys.size();
  }
  out->total_keys = (int64_t)keys.size();
  out->cat_offsets = dupv(offs);
  out->cat_keys = dupv(keys);
  out->cat_counts = dupv(counts);
  if (c->kind == CFB_TRIPLE) {
    std::vector<double> nc((size_t)n * keys.size());
    for (int i = 0; i < n; i++)
      for (size_t t = 0; t < keys.size(); t++)
        nc[(size_t)i * keys.size() + t] = f[L.numcat_base + (long long)i * L.total_dom + dense_t[t]];
    out->numcat_sums = dupv(nc);
    // sparse pair counts: pull this group's hash partition and order it like std::map iteration
    struct HashedPair {
      int pair;
      long long sk, sl;
      unsigned long long count;
    };
    std::vector<HashedPair> hashed;
    if (L.pairs_hashed && c->hash.capacity) {
      const size_t cap = (size_t)c->hash.capacity;
      std::vector<unsigned long long> hk(cap), hc(cap);
      CU(cudaMemcpyAsync(hk.data(), c->hash.keys + (size_t)group * cap, cap * 8, cudaMemcpyDeviceToHost, c->stream));
      CU(cudaMemcpyAsync(hc.data(), c->hash.counts + (size_t)group * cap, cap * 8, cudaMemcpyDeviceToHost, c->stream));
      CU(cudaStreamSynchronize(c->stream));
      const unsigned long long mask = (1ull << cfb::kPairSlotBits) - 1;
      for (size_t i = 0; i < cap; i++)
        if (hk[i] != cfb::kPairEmpty && hc[i])
          hashed.push_back({(int)(hk[i] >> (2 * cfb::kPairSlotBits)), (long long)((hk[i] >> cfb::kPairSlotBits) & mask),
                            (long long)(hk[i] & mask), hc[i]});
      std::sort(hashed.begin(), hashed.end(), [](const HashedPair &a, const HashedPair &b) {
        if (a.pair != b.pair) return a.pair < b.pair;
        if (a.sk != b.sk) return a.sk < b.sk;
        return a.sl < b.sl;
      });
    }
    out->n_pair_lists = (int64_t)m * (m + 1) / 2;
    std::vector<int64_t> po(out->n_pair_lists + 1, 0), pc;
    std::vector<int32_t> k1, k2;
    int p = 0;
    for (int k = 0; k < m; k++)
      for (int l = k; l < m; l++, p++) {
        if (k == l) {
          // (key,key,count): the diagonal pair table is the key count itself
          for (int64_t t = offs[k]; t < offs[k + 1]; t++) {
            k1.push_back(keys[t]);
            k2.push_back(keys[t]);
            pc.push_back(counts[t]);
          }
        } else if (L.pairs_hashed) {
          auto range = std::equal_range(hashed.begin(), hashed.end(), HashedPair{k * m + l, 0, 0, 0},
                                        [](const HashedPair &a, const HashedPair &b) { return a.pair < b.pair; });
          std::vector<std::tuple<int32_t, int32_t, int64_t>> lst;
          for (auto it = range.first; it != range.second; ++it)
            lst.emplace_back(key_of(k, it->sk), key_of(l, it->sl), (int64_t)it->count);
          if (c->dict[k].on || c->dict[l].on) std::sort(lst.begin(), lst.end());
          for (const auto &e : lst) {
            k1.push_back(std::get<0>(e));
            k2.push_back(std::get<1>(e));
            pc.push_back(std::get<2>(e));
          }
        } else {
          const unsigned long long *tab = u.data() + L.pair_base + L.pair_off[k * m + l];
          // only slots whose keys occurred can be non-zero: walk the occurred keys of k and l
          for (int64_t a = offs[k]; a < offs[k + 1]; a++) {
            const long long sk = dense_t[a] - L.cat_off[k];
            for (int64_t b = offs[l]; b < offs[l + 1]; b++) {
              const long long s2 = dense_t[b] - L.cat_off[l];
              const unsigned long long cnt = tab[sk * L.dom[l] + s2];
              if (cnt) {
                k1.push_back(keys[a]);
                k2.push_back(keys[b]);
                pc.push_back((int64_t)cnt);
              }
            }
          }
        }
        po[p + 1] = (int64_t)k1.size();
      }
    out->pair_offsets = dupv(po);
    out->pair_key1 = dupv(k1);
    out->pair_key2 = dupv(k2);
    out->pair_counts = dupv(pc);
  } else {
    out->n_pair_lists = 0;
    out->pair_offsets = dupv(std::vector<int64_t>(1, 0));
  }
  return CFB_OK;
}

int cfb_result_multiply(const cfb_result *a, const cfb_result *b, cfb_result *out) {
  NvtxRange nvtx_range("cfb_result_multiply");
  if (!a || !b || !out) return fail(CFB_ERR_INVALID, "NULL argument");
  if (a->kind != b->kind) return fail(CFB_ERR_INVALID, "multiply: a triple and an NB aggregate do not mix");
  const bool nb = a->kind == CFB_NB;
  const int na = a->n_num, nbn = b->n_num, n = na + nbn, ma = a->n_cat, mb = b->n_cat, m = ma + mb;
  if (n > CFB_MAX_NUM || m > CFB_MAX_CAT) return fail(CFB_ERR_INVALID, "multiply: too many columns in the product");
  memset(out, 0, sizeof(*out));
  const double Na = (double)a->N, Nb = (double)b->N;
  auto dupv = [](const auto &v) {
    using T = typename std::decay<decltype(v)>::type::value_type;
    T *p = (T *)malloc(std::max<size_t>(1, v.size()) * sizeof(T));
    if (!v.empty()) memcpy(p, v.data(), v.size() * sizeof(T));
    return p;
  };
  out->kind = a->kind;
  out->n_num = n;
  out->n_cat = m;
  out->N = a->N * b->N;
  out->n_quad = nb ? n : (int64_t)n * (n + 1) / 2;
  std::vector<double> lin(n), quad(out->n_quad, 0.0);
  for (int i = 0; i < na; i++) lin[i] = Nb * a->lin[i];
  for (int i = 0; i < nbn; i++) lin[na + i] = Na * b->lin[i];
  if (nb) {
    for (int i = 0; i < na; i++) quad[i] = Nb * a->quad[i];
    for (int i = 0; i < nbn; i++) quad[na + i] = Na * b->quad[i];
  } else {
    auto tri = [](int nn, int i, int j) { return i * nn - i * (i + 1) / 2 + j; };
    for (int i = 0; i < n; i++)
      for (int j = i; j < n; j++) {
        double v;
        if (j < na)
          v = Nb * a->quad[tri(na, i, j)];
        else if (i >= na)
          v = Na * b->quad[tri(nbn, i - na, j - na)];
        else
          v = a->lin[i] * b->lin[j - na];
        quad[tri(n, i, j)] = v;
      }
  }
  out->lin = dupv(lin);
  out->quad = dupv(quad);
  // categorical columns: a's, then b's
  const int64_t ta = a->total_keys, tb = b->total_keys, tk = ta + tb;
  std::vector<int64_t> offs(m + 1), counts(tk);
  std::vector<int32_t> keys(tk);
  for (int c = 0; c <= ma; c++) offs[c] = a->cat_offsets[c];
  for (int c = 0; c <= mb; c++) offs[ma + c] = ta + b->cat_offsets[c];
  for (int64_t t = 0; t < ta; t++) {
    keys[t] = a->cat_keys[t];
    counts[t] = a->cat_counts[t] * b->N;
  }
  for (int64_t t = 0; t < tb; t++) {
    keys[ta + t] = b->cat_keys[t];
    counts[ta + t] = b->cat_counts[t] * a->N;
  }
  out->total_keys = tk;
  out->cat_offsets = dupv(offs);
  out->cat_keys = dupv(keys);
  out->cat_counts = dupv(counts);
  if (nb) {
    out->n_pair_lists = 0;
    out->pair_offsets = dupv(std::vector<int64_t>(1, 0));
    return CFB_OK;
  }
  std::vector<double> nc((size_t)n * tk);
  for (int i = 0; i < n; i++)
    for (int64_t t = 0; t < tk; t++) {
      double v;
      if (i < na)
        v = t < ta ? Nb * a->numcat_sums[(size_t)i * ta + t] : a->lin[i] * (double)b->cat_counts[t - ta];
      else
        v = t < ta ? b->lin[i - na] * (double)a->cat_counts[t] : Na * b->numcat_sums[(size_t)(i - na) * tb + (t - ta)];
      nc[(size_t)i * tk + t] = v;
    }
  out->numcat_sums = dupv(nc);
  out->n_pair_lists = (int64_t)m * (m + 1) / 2;
  std::vector<int64_t> po(out->n_pair_lists + 1, 0), pc;
  std::vector<int32_t> k1, k2;
  auto pair_index = [](int mm, int k, int l) { return k * mm - k * (k + 1) / 2 + l; };
  int p = 0;
  for (int k = 0; k < m; k++)
    for (int l = k; l < m; l++, p++) {
      if (l < ma) {  // both on a's side
        const int q = pair_index(ma, k, l);
        for (int64_t t = a->pair_offsets[q]; t < a->pair_offsets[q + 1]; t++) {
          k1.push_back(a->pair_key1[t]);
          k2.push_back(a->pair_key2[t]);
          pc.push_back(a->pair_counts[t] * b->N);
        }
      } else if (k >= ma) {  // both on b's side
        const int q = pair_index(mb, k - ma, l - ma);
        for (int64_t t = b->pair_offsets[q]; t < b->pair_offsets[q + 1]; t++) {
          k1.push_back(b->pair_key1[t]);
          k2.push_back(b->pair_key2[t]);
          pc.push_back(b->pair_counts[t] * a->N);
        }
      } else {  // across the join: outer product of the key counts (ascending (key1,key2))
        for (int64_t x = a->cat_offsets[k]; x < a->cat_offsets[k + 1]; x++)
          for (int64_t y = b->cat_offsets[l - ma]; y < b->cat_offsets[l - ma + 1]; y++) {
            k1.push_back(a->cat_keys[x]);
            k2.push_back(b->cat_keys[y]);
            pc.push_back(a->cat_counts[x] * b->cat_counts[y]);
          }
      }
      po[p + 1] = (int64_t)k1.size();
    }
  out->pair_offsets = dupv(po);
  out->pair_key1 = dupv(k1);
  out->pair_key2 = dupv(k2);
  out->pair_counts = dupv(pc);
  return CFB_OK;
}

int cfb_result_combine(const cfb_result *a, const cfb_result *b, int sign, int flags, cfb_result *out) {
  NvtxRange nvtx_range("cfb_result_combine");
  if (!a || !b || !out) return fail(CFB_ERR_INVALID, "NULL argument");
  if (sign != 1 && sign != -1) return fail(CFB_ERR_INVALID, "sign must be +1 or -1");
  if (flags & ~CFB_COMBINE_KEEP_ZERO_KEYS) return fail(CFB_ERR_INVALID, "combine: unknown flag");
  const bool keep_zero = (flags & CFB_COMBINE_KEEP_ZERO_KEYS) != 0;
  if (a->kind != b->kind || a->n_num != b->n_num || a->n_cat != b->n_cat || a->n_quad != b->n_quad)
    return fail(CFB_ERR_INVALID, "combine: the two results have different shapes");
  const bool nb = a->kind == CFB_NB;
  const int n = a->n_num, m = a->n_cat;
  const double sg = (double)sign;
  memset(out, 0, sizeof(*out));
  auto dupv = [](const auto &v) {
    using T = typename std::decay<decltype(v)>::type::value_type;
    T *p = (T *)malloc(std::max<size_t>(1, v.size()) * sizeof(T));
    if (!v.empty()) memcpy(p, v.data(), v.size() * sizeof(T));
    return p;
  };
  out->kind = a->kind;
  out->n_num = n;
  out->n_cat = m;
  out->N = a->N + sign * b->N;
  out->n_quad = a->n_quad;
  std::vector<double> lin(n), quad(a->n_quad);
  for (int i = 0; i < n; i++) lin[i] = a->lin[i] + sg * b->lin[i];
  for (int64_t i = 0; i < a->n_quad; i++) quad[i] = a->quad[i] + sg * b->quad[i];
  out->lin = dupv(lin);
  out->quad = dupv(quad);
  // categorical columns: merge the ascending key lists, drop keys whose count becomes 0
  std::vector<int64_t> offs(m + 1, 0), counts;
  std::vector<int32_t> keys;
  std::vector<std::pair<int64_t, int64_t>> src;  // per output key: its position in a / in b (-1 = absent)
  for (int c = 0; c < m; c++) {
    int64_t x = a->cat_offsets[c], xe = a->cat_offsets[c + 1], y = b->cat_offsets[c], ye = b->cat_offsets[c + 1];
    while (x < xe || y < ye) {
      int64_t px = -1, py = -1;
      int32_t key;
      if (y >= ye || (x < xe && a->cat_keys[x] < b->cat_keys[y])) key = a->cat_keys[px = x++];
      else if (x >= xe || b->cat_keys[y] < a->cat_keys[x]) key = b->cat_keys[py = y++];
      else {
        key = a->cat_keys[x];
        px = x++;
        py = y++;
      }
      const int64_t cnt = (px >= 0 ? a->cat_counts[px] : 0) + sign * (py >= 0 ? b->cat_counts[py] : 0);
      if (cnt == 0 && !keep_zero) continue;
      keys.push_back(key);
      counts.push_back(cnt);
      src.emplace_back(px, py);
    }
    offs[c + 1] = (int64_t)keys.size();
  }
  const int64_t tk = (int64_t)keys.size();
  out->total_keys = tk;
  out->cat_offsets = dupv(offs);
  out->cat_keys = dupv(keys);
  out->cat_counts = dupv(counts);
  if (nb) {
    out->n_pair_lists = 0;
    out->pair_offsets = dupv(std::vector<int64_t>(1, 0));
    return CFB_OK;
  }
  std::vector<double> nc((size_t)n * tk);
  for (int i = 0; i < n; i++)
    for (int64_t t = 0; t < tk; t++)
      nc[(size_t)i * tk + t] = (src[t].first >= 0 ? a->numcat_sums[(size_t)i * a->total_keys + src[t].first] : 0.0) +
                               sg * (src[t].second >= 0 ? b->numcat_sums[(size_t)i * b->total_keys + src[t].second] : 0.0);
  out->numcat_sums = dupv(nc);
  out->n_pair_lists = a->n_pair_lists;
  std::vector<int64_t> po(a->n_pair_lists + 1, 0), pc;
  std::vector<int32_t> k1, k2;
  for (int64_t p = 0; p < a->n_pair_lists; p++) {
    int64_t x = a->pair_offsets[p], xe = a->pair_offsets[p + 1], y = b->pair_offsets[p], ye = b->pair_offsets[p + 1];
    auto less = [](int32_t a1, int32_t a2, int32_t b1, int32_t b2) { return a1 < b1 || (a1 == b1 && a2 < b2); };
    while (x < xe || y < ye) {
      int64_t cnt;
      int32_t u, v;
      if (y >= ye || (x < xe && less(a->pair_key1[x], a->pair_key2[x], b->pair_key1[y], b->pair_key2[y]))) {
        u = a->pair_key1[x], v = a->pair_key2[x], cnt = a->pair_counts[x++];
      } else if (x >= xe || less(b->pair_key1[y], b->pair_key2[y], a->pair_key1[x], a->pair_key2[x])) {
        u = b->pair_key1[y], v = b->pair_key2[y], cnt = sign * b->pair_counts[y++];
      } else {
        u = a->pair_key1[x], v = a->pair_key2[x], cnt = a->pair_counts[x++] + sign * b->pair_counts[y++];
      }
      if (cnt == 0 && !keep_zero) continue;
      k1.push_back(u);
      k2.push_back(v);
      pc.push_back(cnt);
    }
    po[p + 1] = (int64_t)k1.size();
  }
  out->pair_offsets = dupv(po);
  out->pair_key1 = dupv(k1);
  out->pair_key2 = dupv(k2);
  out->pair_counts = dupv(pc);
  return CFB_OK;
}

int cfb_result_impute_linear(const cfb_result *a, const cfb_linear_model *M, int target, cfb_result *out) {
  NvtxRange nvtx_range("cfb_result_impute_linear");
  if (!a || !M || !out) return fail(CFB_ERR_INVALID, "NULL argument");
  if (a->kind != CFB_TRIPLE) return fail(CFB_ERR_INVALID, "impute_linear needs the full ring (CFB_TRIPLE)");
  const int n = a->n_num, m = a->n_cat;
  if (target < 0 || target >= n) return fail(CFB_ERR_INVALID, "target %d is not a numeric column", target);
  if (M->n_out != 1 || M->n_num != n - 1 || M->n_cat != m)
    return fail(CFB_ERR_INVALID, "model shape (%d numeric, %d categorical, %d outputs) does not fit the cofactor without its target",
                M->n_num, M->n_cat, M->n_out);
  const int64_t tk = a->total_keys;
  // theta over [1 | numeric (0 for the target) | result keys]: a key the model does not hold weighs 0 (as in predict)
  const double b = M->bias[0];
  std::vector<double> wn(n, 0.0), wk((size_t)tk, 0.0);
  for (int i = 0, f = 0; i < n; i++)
    if (i != target) wn[i] = M->w_num[f++];
  for (int c = 0; c < m; c++) {
    const int32_t *mk = M->cat_keys + M->cat_offsets[c], *me = M->cat_keys + M->cat_offsets[c + 1];
    for (int64_t t = a->cat_offsets[c]; t < a->cat_offsets[c + 1]; t++) {
      const int32_t *hit = std::lower_bound(mk, me, a->cat_keys[t]);
      if (hit != me && *hit == a->cat_keys[t]) wk[(size_t)t] = M->w_cat[M->cat_offsets[c] + (hit - mk)];
    }
  }
  auto quad_at = [&](int i, int j) -> double & {
    if (i > j) std::swap(i, j);
    return a->quad[(int64_t)i * n - (int64_t)i * (i + 1) / 2 + j];
  };
  // copy, then rewrite what involves the target
  {  // a + 0 with the keys kept: a deep copy through the combine path
    cfb_result z = *a;
    std::vector<double> zl(n, 0.0), zq((size_t)a->n_quad, 0.0), zn((size_t)n * tk, 0.0);
    std::vector<int64_t> zc((size_t)tk, 0), zp((size_t)(a->n_pair_lists ? a->pair_offsets[a->n_pair_lists] : 0), 0);
    z.N = 0;
    z.lin = zl.data();
    z.quad = zq.data();
    z.numcat_sums = zn.data();
    z.cat_counts = zc.data();
    z.pair_counts = zp.data();
    int rc = cfb_result_combine(a, &z, +1, CFB_COMBINE_KEEP_ZERO_KEYS, out);
    if (rc) return rc;
  }
  // SUM y
  double sy = b * (double)a->N;
  for (int k = 0; k < n; k++) sy += wn[k] * a->lin[k];
  for (int64_t t = 0; t < tk; t++) sy += wk[(size_t)t] * (double)a->cat_counts[t];
  // SUM y x_i
  std::vector<double> syx(n, 0.0);
  for (int i = 0; i < n; i++) {
    if (i == target) continue;
    double v = b * a->lin[i];
    for (int k = 0; k < n; k++)
      if (k != target) v += wn[k] * quad_at(k, i);
    for (int64_t t = 0; t < tk; t++) v += wk[(size_t)t] * a->numcat_sums[(size_t)i * tk + t];
    syx[i] = v;
  }
  // SUM y [key_d = kappa]
  std::vector<double> syk((size_t)tk, 0.0);
  for (int64_t t = 0; t < tk; t++) {
    double v = b * (double)a->cat_counts[t];
    for (int k = 0; k < n; k++)
      if (k != target) v += wn[k] * a->numcat_sums[(size_t)k * tk + t];
    v += wk[(size_t)t] * (double)a->cat_counts[t];  // the same column: the key meets only itself
    syk[(size_t)t] = v;
  }
  auto entry_of = [&](int c, int32_t key) -> int64_t {
    const int32_t *lo = a->cat_keys + a->cat_offsets[c], *hi = a->cat_keys + a->cat_offsets[c + 1];
    const int32_t *hit = std::lower_bound(lo, hi, key);
    return hit != hi && *hit == key ? a->cat_offsets[c] + (hit - lo) : -1;
  };
  int64_t list = 0;
  for (int k = 0; k < m; k++)
    for (int l = k; l < m; l++, list++) {
      if (k == l) continue;
      for (int64_t t = a->pair_offsets[list]; t < a->pair_offsets[list + 1]; t++) {
        const int64_t ek = entry_of(k, a->pair_key1[t]), el = entry_of(l, a->pair_key2[t]);
        if (ek < 0 || el < 0) {
          cfb_result_free(out);
          return fail(CFB_ERR_INVALID, "impute_linear: a key pair names a key its column's list does not hold");
        }
        const double cnt = (double)a->pair_counts[t];
        syk[(size_t)el] += wk[(size_t)ek] * cnt;
        syk[(size_t)ek] += wk[(size_t)el] * cnt;
      }
    }
  // SUM y^2 = theta . (the new cross terms)
  double syy = b * sy;
  for (int k = 0; k < n; k++)
    if (k != target) syy += wn[k] * syx[k];
  for (int64_t t = 0; t < tk; t++) syy += wk[(size_t)t] * syk[(size_t)t];
  out->lin[target] = sy;
  for (int i = 0; i < n; i++) {
    const int lo = std::min(i, target), hi = std::max(i, target);
    out->quad[(int64_t)lo * n - (int64_t)lo * (lo + 1) / 2 + hi] = i == target ? syy : syx[i];
  }
  for (int64_t t = 0; t < tk; t++) out->numcat_sums[(size_t)target * tk + t] = syk[(size_t)t];
  return CFB_OK;
}

void cfb_result_free(cfb_result *r) {
  if (!r) return;
  free(r->lin);
  free(r->quad);
  free(r->cat_offsets);
  free(r->cat_keys);
  free(r->cat_counts);
  free(r->numcat_sums);
  free(r->pair_offsets);
  free(r->pair_key1);
  free(r->pair_key2);
  free(r->pair_counts);
  memset(r, 0, sizeof(*r));
}

int cfb_ctx_partial_sizes(cfb_ctx *c, size_t *n_f64, size_t *n_u64) {
  if (!c || !n_f64 || !n_u64) return fail(CFB_ERR_INVALID, "NULL argument");
  *n_f64 = (size_t)(c->lay.F * c->lay.n_groups);
  *n_u64 = (size_t)(c->lay.U * c->lay.n_groups);
  return CFB_OK;
}

int cfb_ctx_export_partial(cfb_ctx *c, void *d_f64, void *d_u64, void *stream) {
  if (!c || !d_f64 || !d_u64) return fail(CFB_ERR_INVALID, "NULL argument");
  if (c->lay.pairs_hashed || any_dict(c))
    return fail(CFB_ERR_DOMAIN, "sparse state (hashed pair counts / key dictionaries) has no dense partial: combine with cfb_ctx_combine");
  CU(cudaSetDevice(c->device));
  int rc = flush_tile(c);
  if (rc) return rc;
  // Stream-ordered when the caller names the stream its scans and its collective run on (no host
  // synchronisation between scan, export, NCCL and import); synchronous on the context's own stream.
  cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
  if (stream && c->fill == 0 && c->tile_rows) CU(cudaStreamSynchronize(c->stream));  // staged tiles ran on the context stream
  if (c->lay.F) CU(cudaMemcpyAsync(d_f64, c->d_f64, c->lay.F * c->lay.n_groups * 8, cudaMemcpyDeviceToDevice, s));
  CU(cudaMemcpyAsync(d_u64, c->d_u64, c->lay.U * c->lay.n_groups * 8, cudaMemcpyDeviceToDevice, s));
  if (!stream) CU(cudaStreamSynchronize(s));
  return CFB_OK;
}

int cfb_ctx_import_partial(cfb_ctx *c, const void *d_f64, const void *d_u64, void *stream) {
  if (!c || !d_f64 || !d_u64) return fail(CFB_ERR_INVALID, "NULL argument");
  if (c->lay.pairs_hashed || any_dict(c))
    return fail(CFB_ERR_DOMAIN, "sparse state (hashed pair counts / key dictionaries) has no dense partial: combine with cfb_ctx_combine");
  CU(cudaSetDevice(c->device));
  int rc = flush_tile(c);
  if (rc) return rc;
  c->touched = true;
  cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
  if (stream) {
    c->user_stream = s;  // finalize / sync wait for it
    if (c->tile_rows) CU(cudaStreamSynchronize(c->stream));
  }
  if (c->lay.F) CU(cudaMemcpyAsync(c->d_f64, d_f64, c->lay.F * c->lay.n_groups * 8, cudaMemcpyDeviceToDevice, s));
  CU(cudaMemcpyAsync(c->d_u64, d_u64, c->lay.U * c->lay.n_groups * 8, cudaMemcpyDeviceToDevice, s));
  if (!stream) CU(cudaStreamSynchronize(s));
  return CFB_OK;
}

int cfb_cat_minmax_device(int device, const int32_t *const *d_cat_cols, int n_cat, size_t n_rows, int32_t *lo_out,
                          int32_t *hi_out, void *stream) {
  if (!d_cat_cols || !lo_out || !hi_out || n_cat < 0 || n_cat > CFB_MAX_CAT) return fail(CFB_ERR_INVALID, "bad argument");
  if (device_count_quiet() == 0) return fail(CFB_ERR_NO_DEVICE, "no CUDA device is visible");
  if (n_cat == 0) return CFB_OK;
  CU(cudaSetDevice(device));
  cudaStream_t s = (cudaStream_t)stream;
  int h[2 * cfb::kMaxCat];
  for (int k = 0; k < cfb::kMaxCat; k++) {
    h[k] = INT_MAX;
    h[cfb::kMaxCat + k] = INT_MIN;
  }
  int *d = nullptr;
  CU(cudaMalloc(&d, sizeof(h)));
  CU(cudaMemcpyAsync(d, h, sizeof(h), cudaMemcpyHostToDevice, s));
  cfb::ScanCols sc{};
  for (int k = 0; k < n_cat; k++) sc.cat[k] = d_cat_cols[k];
  if (n_rows) {
    const int bx = (int)std::min<size_t>((n_rows + 255) / 256, (size_t)std::max(1, dev_info(device).sms * 8 / n_cat));
    cfb::cat_minmax_kernel<<<dim3(std::max(bx, 1), n_cat), 256, 0, s>>>(sc, n_rows, d, d + cfb::kMaxCat);
    g_launches++;
  }
  CU(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  cudaFree(d);
  for (int k = 0; k < n_cat; k++) {
    if (h[k] > h[cfb::kMaxCat + k]) h[k] = h[cfb::kMaxCat + k] = 0;  // no rows
    lo_out[k] = h[k];
    hi_out[k] = h[cfb::kMaxCat + k];
  }
  return CFB_OK;
}

// ------------------------------------------------------- synthetic data (tests, bench)
int cfb_gen_uniform_f32(int device, float *d_out, size_t n, uint64_t seed, uint64_t first, void *stream) {
  if (device_count_quiet() == 0) return fail(CFB_ERR_NO_DEVICE, "no CUDA device is visible");
  CU(cudaSetDevice(device));
  if (!n) return CFB_OK;
  const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)dev_info(device).sms * 16);
  cfb::gen_uniform_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_out, n, seed, first);
  g_launches++;
  CU(cudaGetLastError());
  return CFB_OK;
}

int cfb_gen_int32(int device, int32_t *d_out, size_t n, uint64_t seed, uint64_t first, int32_t lo, uint32_t range,
                  void *stream) {
  if (device_count_quiet() == 0) return fail(CFB_ERR_NO_DEVICE, "no CUDA device is visible");
  if (range == 0) return fail(CFB_ERR_INVALID, "range must be > 0");
  CU(cudaSetDevice(device));
  if (!n) return CFB_OK;
  const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)dev_info(device).sms * 16);
  cfb::gen_int_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_out, n, seed, first, lo, range);
  g_launches++;
  CU(cudaGetLastError());
  return CFB_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ multi-GPU exchange (NCCL)
// The library does not link NCCL: the functions are resolved from libnccl.so.2 at first use, so a host that
// never reduces across GPUs needs no NCCL at all, and a host that already loaded one (PyTorch ships its own)
// shares that copy.
namespace {
struct NcclApi {
  void *handle = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommCount)(const ncclComm_t, int *) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  std::string error;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;

const NcclApi &nccl_api() {
  std::call_once(g_nccl_once, [] {
    NcclApi &a = g_nccl;
    const char *names[] = {getenv("CFB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
      if (!nm || !*nm) continue;
      a.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (a.handle) break;
    }
    if (!a.handle) {
      a.error = std::string("libnccl.so.2 could not be loaded: ") + (dlerror() ? dlerror() : "?");
      return;
    }
    auto sym = [&](const char *n) {
      void *p = dlsym(a.handle, n);
      if (!p && a.error.empty()) a.error = std::string("NCCL symbol missing: ") + n;
      return p;
    };
    a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
    a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
    a.CommCount = (decltype(a.CommCount))sym("ncclCommCount");
    a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
    a.AllGather = (decltype(a.AllGather))sym("ncclAllGather");
    a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
  });
  return g_nccl;
}

#define NC(call)                                                                                              \
  do {                                                                                                        \
    ncclResult_t r_ = (call);                                                                                 \
    if (r_ != ncclSuccess) return fail(CFB_ERR_CUDA, "%s failed: %s", #call, nccl_api().GetErrorString(r_)); \
  } while (0)

int nccl_ready() {
  const NcclApi &a = nccl_api();
  if (!a.error.empty()) return fail(CFB_ERR_STATE, "%s", a.error.c_str());
  return CFB_OK;
}
}  // namespace

extern "C" {

int cfb_nccl_unique_id(void *id128) {
  if (!id128) return fail(CFB_ERR_INVALID, "id buffer is NULL");
  int rc = nccl_ready();
  if (rc) return rc;
  static_assert(sizeof(ncclUniqueId) == CFB_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  NC(nccl_api().GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return CFB_OK;
}

int cfb_nccl_comm_create(int device, int world, int rank, const void *id128, void **comm_out) {
  if (!id128 || !comm_out) return fail(CFB_ERR_INVALID, "NULL argument");
  *comm_out = nullptr;
  if (world < 1 || rank < 0 || rank >= world) return fail(CFB_ERR_INVALID, "bad rank %d / world %d", rank, world);
  if (device_count_quiet() == 0) return fail(CFB_ERR_NO_DEVICE, "no CUDA device is visible");
  int rc = nccl_ready();
  if (rc) return rc;
  CU(cudaSetDevice(device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm = nullptr;
  NC(nccl_api().CommInitRank(&comm, world, id, rank));
  *comm_out = comm;
  return CFB_OK;
}

int cfb_nccl_comm_destroy(void *comm) {
  if (!comm) return CFB_OK;
  int rc = nccl_ready();
  if (rc) return rc;
  NC(nccl_api().CommDestroy((ncclComm_t)comm));
  return CFB_OK;
}

int cfb_nccl_agree_domain(void *comm, int device, int32_t *lo, int32_t *hi, int n_cat, void *stream) {
  if (!comm || !lo || !hi || n_cat < 0 || n_cat > CFB_MAX_CAT) return fail(CFB_ERR_INVALID, "bad argument");
  if (n_cat == 0) return CFB_OK;
  int rc = nccl_ready();
  if (rc) return rc;
  CU(cudaSetDevice(device));
  cudaStream_t s = (cudaStream_t)stream;
  int32_t *d = nullptr;
  CU(cudaMalloc(&d, 2 * CFB_MAX_CAT * sizeof(int32_t)));
  CU(cudaMemcpyAsync(d, lo, n_cat * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(d + CFB_MAX_CAT, hi, n_cat * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  const NcclApi &a = nccl_api();
  ncclResult_t r = a.GroupStart();
  if (r == ncclSuccess) r = a.AllReduce(d, d, n_cat, ncclInt32, ncclMin, (ncclComm_t)comm, s);
  if (r == ncclSuccess) r = a.AllReduce(d + CFB_MAX_CAT, d + CFB_MAX_CAT, n_cat, ncclInt32, ncclMax, (ncclComm_t)comm, s);
  const ncclResult_t r2 = a.GroupEnd();
  if (r == ncclSuccess) r = r2;
  if (r != ncclSuccess) {
    cudaFree(d);
    return fail(CFB_ERR_CUDA, "NCCL min/max all-reduce failed: %s", a.GetErrorString(r));
  }
  CU(cudaMemcpyAsync(lo, d, n_cat * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(hi, d + CFB_MAX_CAT, n_cat * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  cudaFree(d);
  return CFB_OK;
}

// A result as one flat byte string (and back): what ranks exchange when the state has no dense partial.
static void pack_result(const cfb_result &r, std::vector<char> &out) {
  auto put = [&](const void *p, size_t bytes) {
    const char *b = (const char *)p;
    out.insert(out.end(), b, b + bytes);
  };
  const int64_t np = r.n_pair_lists ? r.pair_offsets[r.n_pair_lists] : 0;
  const int64_t head[8] = {r.kind, r.n_num, r.n_cat, r.N, r.n_quad, r.total_keys, r.n_pair_lists, np};
  put(head, sizeof head);
  put(r.lin, (size_t)r.n_num * 8);
  put(r.quad, (size_t)r.n_quad * 8);
  put(r.cat_offsets, ((size_t)r.n_cat + 1) * 8);
  put(r.cat_counts, (size_t)r.total_keys * 8);
  if (r.kind == CFB_TRIPLE) put(r.numcat_sums, (size_t)r.n_num * r.total_keys * 8);
  put(r.pair_offsets, ((size_t)r.n_pair_lists + 1) * 8);
  put(r.pair_counts, (size_t)np * 8);
  put(r.cat_keys, (size_t)r.total_keys * 4);
  put(r.pair_key1, (size_t)np * 4);
  put(r.pair_key2, (size_t)np * 4);
  while (out.size() % 8) out.push_back(0);
}
// a VIEW into the byte string (8-byte fields first, so every array is aligned); returns the bytes consumed
static size_t view_result(const char *p, cfb_result *r) {
  const char *p0 = p;
  int64_t head[8];
  memcpy(head, p, sizeof head);
  p += sizeof head;
  memset(r, 0, sizeof *r);
  r->kind = (int32_t)head[0];
  r->n_num = (int32_t)head[1];
  r->n_cat = (int32_t)head[2];
  r->N = head[3];
  r->n_quad = head[4];
  r->total_keys = head[5];
  r->n_pair_lists = head[6];
  const int64_t np = head[7];
  auto take = [&](size_t bytes) {
    const char *q = p;
    p += bytes;
    return (void *)q;
  };
  r->lin = (double *)take((size_t)r->n_num * 8);
  r->quad = (double *)take((size_t)r->n_quad * 8);
  r->cat_offsets = (int64_t *)take(((size_t)r->n_cat + 1) * 8);
  r->cat_counts = (int64_t *)take((size_t)r->total_keys * 8);
  if (r->kind == CFB_TRIPLE) r->numcat_sums = (double *)take((size_t)r->n_num * r->total_keys * 8);
  r->pair_offsets = (int64_t *)take(((size_t)r->n_pair_lists + 1) * 8);
  r->pair_counts = (int64_t *)take((size_t)np * 8);
  r->cat_keys = (int32_t *)take((size_t)r->total_keys * 4);
  r->pair_key1 = (int32_t *)take((size_t)np * 4);
  r->pair_key2 = (int32_t *)take((size_t)np * 4);
  size_t used = (size_t)(p - p0);
  return (used + 7) & ~(size_t)7;
}

// The exchange for states without a dense partial ("sparse / large-domain fallback"): every rank finalizes
// its groups, the canonical results are all-gathered as byte strings (sizes agreed by a MAX all-reduce) and merged by
// key in rank order on every rank -- the same sums in the same order everywhere.  The merged results are kept in the
// context; cfb_ctx_finalize hands them out.
static int allreduce_results(cfb_ctx *c, ncclComm_t comm, cudaStream_t s) {
  const NcclApi &a = nccl_api();
  int world = 1;
  if (a.CommCount(comm, &world) != ncclSuccess) return fail(CFB_ERR_CUDA, "ncclCommCount failed");
  std::vector<char> mine(8, 0);  // [0..8): the length of this rank's string
  for (int g = 0; g < c->G; g++) {
    cfb_result r;
    int rc = cfb_ctx_finalize(c, g, &r);
    if (rc) return rc;
    pack_result(r, mine);
    cfb_result_free(&r);
  }
  const unsigned long long len = mine.size();
  memcpy(mine.data(), &len, 8);
  unsigned long long *d_len = nullptr;
  CU(cudaMalloc(&d_len, 8));
  CU(cudaMemcpyAsync(d_len, &len, 8, cudaMemcpyHostToDevice, s));
  ncclResult_t r = a.AllReduce(d_len, d_len, 1, ncclUint64, ncclMax, comm, s);
  unsigned long long padded = 0;
  cudaError_t e = cudaMemcpyAsync(&padded, d_len, 8, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(d_len);
  if (r != ncclSuccess) return fail(CFB_ERR_CUDA, "NCCL all-reduce of the result sizes failed: %s", a.GetErrorString(r));
  CU(e);
  char *d_send = nullptr, *d_recv = nullptr;
  CU(cudaMalloc(&d_send, padded));
  e = cudaMalloc(&d_recv, padded * world);
  if (e != cudaSuccess) {
    cudaFree(d_send);
    CU(e);
  }
  std::vector<char> all((size_t)padded * world);
  e = cudaMemcpyAsync(d_send, mine.data(), mine.size(), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) r = a.AllGather(d_send, d_recv, padded, ncclChar, comm, s);
  if (e == cudaSuccess && r == ncclSuccess) e = cudaMemcpyAsync(all.data(), d_recv, all.size(), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(d_send);
  cudaFree(d_recv);
  if (r != ncclSuccess) return fail(CFB_ERR_CUDA, "NCCL all-gather of the results failed: %s", a.GetErrorString(r));
  CU(e);
  std::vector<cfb_result> merged((size_t)c->G);
  for (auto &m : merged) memset(&m, 0, sizeof m);
  for (int rank = 0; rank < world; rank++) {
    const char *p = all.data() + (size_t)rank * padded + 8;
    for (int g = 0; g < c->G; g++) {
      cfb_result v;
      p += view_result(p, &v);
      if (rank == 0) {
        copy_result(&v, &merged[(size_t)g]);
        continue;
      }
      cfb_result sum;
      const int rc = cfb_result_combine(&merged[(size_t)g], &v, +1, 0, &sum);
      if (rc) {
        for (auto &m : merged) cfb_result_free(&m);
        return rc;
      }
      cfb_result_free(&merged[(size_t)g]);
      merged[(size_t)g] = sum;
    }
  }
  drop_reduced(c);
  c->reduced = std::move(merged);
  return CFB_OK;
}

int cfb_ctx_allreduce(cfb_ctx *c, void *comm, void *stream) {
  NvtxRange nvtx_range("cfb_ctx_allreduce");
  if (!c || !comm) return fail(CFB_ERR_INVALID, "NULL argument");
  if (int rc0 = check_open(c)) return rc0;
  if (c->lay.pairs_hashed || any_dict(c) || (c->m > 0 && !c->user_domain)) {
    // no dense partial that means the same on every rank (hashed pair counts, key dictionaries, or a domain each
    // rank discovered for itself): exchange canonical results instead.  Every rank must be in the same case --
    // declare the domain on all ranks (cfb_nccl_agree_domain + cfb_ctx_set_cat_domain) or on none.
    int rc = nccl_ready();
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    return allreduce_results(c, (ncclComm_t)comm, stream ? (cudaStream_t)stream : c->stream);
  }
  int rc = nccl_ready();
  if (rc) return rc;
  CU(cudaSetDevice(c->device));
  rc = flush_tile(c);
  if (rc) return rc;
  c->touched = true;
  cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
  if (stream) {
    c->user_stream = s;  // finalize / sync wait for it
    if (c->tile_rows) CU(cudaStreamSynchronize(c->stream));  // staged tiles ran on the context stream
  }
  // ONE fused collective, in place on the state: fp64 sums and u64 counts (exact, order-independent) inside one
  // NCCL group -- no export / import copies, no host synchronisation between the scan and the reduce.
  const NcclApi &a = nccl_api();
  const size_t nf = (size_t)(c->lay.F * c->lay.n_groups), nu = (size_t)(c->lay.U * c->lay.n_groups);
  ncclResult_t r = a.GroupStart();
  if (r == ncclSuccess && nf) r = a.AllReduce(c->d_f64, c->d_f64, nf, ncclDouble, ncclSum, (ncclComm_t)comm, s);
  if (r == ncclSuccess) r = a.AllReduce(c->d_u64, c->d_u64, nu, ncclUint64, ncclSum, (ncclComm_t)comm, s);
  const ncclResult_t r2 = a.GroupEnd();
  if (r == ncclSuccess) r = r2;
  if (r != ncclSuccess) return fail(CFB_ERR_CUDA, "NCCL all-reduce of the partial triple failed: %s", a.GetErrorString(r));
  if (!stream) CU(cudaStreamSynchronize(s));
  return CFB_OK;
}

}  // extern "C"

// ------------------------------------------------------- sigma matrix + trainers
#include "sigma_train.cuh"
