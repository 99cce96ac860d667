// stage_copy.cpp -- the host half of the GPU feed: DuckDB vector data -> pinned staging tiles.
//
// The staging tile is written once by the host and next read by the DMA engine, so the copy uses
// non-temporal stores (no read-for-ownership, no cache pollution): per staged byte the host memory system
// sees one read and one write instead of one read, one RFO read and one write-back.  The widest vector unit
// the CPU has is picked once at load time (AVX-512 / AVX2 / SSE2): full 64-byte lines per store keep the
// write-combining buffers from partial flushes, which is what bounded the 16-byte SSE2 loop of round 1
// (about 2 GB/s per worker thread on the 32-core box).  Compiled by the host compiler (not nvcc) so that the
// per-function target attributes and the AVX-512 intrinsics are available.
#include <immintrin.h>

#include <chrono>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace {

using copy_fn = void (*)(char *, const char *, size_t);
using gather_fn = void (*)(uint32_t *, const uint32_t *, const uint32_t *, size_t);

// ---- contiguous copy, dst 64-byte aligned, bytes a multiple of 64 (the callers peel head and tail)
void copy_lines_sse2(char *d, const char *s, size_t bytes) {
  for (size_t i = 0; i < bytes; i += 64) {
    const __m128i a = _mm_loadu_si128((const __m128i *)(s + i)), b = _mm_loadu_si128((const __m128i *)(s + i + 16));
    const __m128i c = _mm_loadu_si128((const __m128i *)(s + i + 32)), e = _mm_loadu_si128((const __m128i *)(s + i + 48));
    _mm_stream_si128((__m128i *)(d + i), a);
    _mm_stream_si128((__m128i *)(d + i + 16), b);
    _mm_stream_si128((__m128i *)(d + i + 32), c);
    _mm_stream_si128((__m128i *)(d + i + 48), e);
  }
}

__attribute__((target("avx2"))) void copy_lines_avx2(char *d, const char *s, size_t bytes) {
  size_t i = 0;
  for (; i + 128 <= bytes; i += 128) {
    const __m256i a = _mm256_loadu_si256((const __m256i *)(s + i)), b = _mm256_loadu_si256((const __m256i *)(s + i + 32));
    const __m256i c = _mm256_loadu_si256((const __m256i *)(s + i + 64)), e = _mm256_loadu_si256((const __m256i *)(s + i + 96));
    _mm256_stream_si256((__m256i *)(d + i), a);
    _mm256_stream_si256((__m256i *)(d + i + 32), b);
    _mm256_stream_si256((__m256i *)(d + i + 64), c);
    _mm256_stream_si256((__m256i *)(d + i + 96), e);
  }
  for (; i < bytes; i += 64) {
    _mm256_stream_si256((__m256i *)(d + i), _mm256_loadu_si256((const __m256i *)(s + i)));
    _mm256_stream_si256((__m256i *)(d + i + 32), _mm256_loadu_si256((const __m256i *)(s + i + 32)));
  }
}

__attribute__((target("avx512f"))) void copy_lines_avx512(char *d, const char *s, size_t bytes) {
  size_t i = 0;
  for (; i + 256 <= bytes; i += 256) {
    const __m512i a = _mm512_loadu_si512(s + i), b = _mm512_loadu_si512(s + i + 64);
    const __m512i c = _mm512_loadu_si512(s + i + 128), e = _mm512_loadu_si512(s + i + 192);
    _mm512_stream_si512((__m512i *)(d + i), a);
    _mm512_stream_si512((__m512i *)(d + i + 64), b);
    _mm512_stream_si512((__m512i *)(d + i + 128), c);
    _mm512_stream_si512((__m512i *)(d + i + 192), e);
  }
  for (; i < bytes; i += 64) _mm512_stream_si512((__m512i *)(d + i), _mm512_loadu_si512(s + i));
}

// ---- dst[i] = src[sel[i]] for 4-byte elements (DICTIONARY / filtered vectors)
void gather_scalar(uint32_t *d, const uint32_t *s, const uint32_t *sel, size_t n) {
  for (size_t i = 0; i < n; i++) d[i] = s[sel[i]];
}
__attribute__((target("avx2"))) void gather_avx2(uint32_t *d, const uint32_t *s, const uint32_t *sel, size_t n) {
  size_t i = 0;
  for (; i + 8 <= n; i += 8) {
    const __m256i idx = _mm256_loadu_si256((const __m256i *)(sel + i));
    _mm256_storeu_si256((__m256i *)(d + i), _mm256_i32gather_epi32((const int *)s, idx, 4));
  }
  for (; i < n; i++) d[i] = s[sel[i]];
}

// ---- key range of a staged categorical vector (domain discovery): the data is in cache, the loop must not be scalar
using minmax_fn = void (*)(const int32_t *, size_t, int32_t *, int32_t *);
void minmax_scalar(const int32_t *s, size_t n, int32_t *lo, int32_t *hi) {
  int32_t a = *lo, b = *hi;
  for (size_t i = 0; i < n; i++) {
    a = s[i] < a ? s[i] : a;
    b = s[i] > b ? s[i] : b;
  }
  *lo = a;
  *hi = b;
}
__attribute__((target("avx2"))) void minmax_avx2(const int32_t *s, size_t n, int32_t *lo, int32_t *hi) {
  size_t i = 0;
  if (n >= 16) {
    __m256i a0 = _mm256_loadu_si256((const __m256i *)s), a1 = _mm256_loadu_si256((const __m256i *)(s + 8)), b0 = a0, b1 = a1;
    for (i = 16; i + 16 <= n; i += 16) {
      const __m256i x = _mm256_loadu_si256((const __m256i *)(s + i)), y = _mm256_loadu_si256((const __m256i *)(s + i + 8));
      a0 = _mm256_min_epi32(a0, x), b0 = _mm256_max_epi32(b0, x);
      a1 = _mm256_min_epi32(a1, y), b1 = _mm256_max_epi32(b1, y);
    }
    alignas(32) int32_t va[8], vb[8];
    _mm256_store_si256((__m256i *)va, _mm256_min_epi32(a0, a1));
    _mm256_store_si256((__m256i *)vb, _mm256_max_epi32(b0, b1));
    minmax_scalar(va, 8, lo, hi);
    minmax_scalar(vb, 8, lo, hi);
  }
  minmax_scalar(s + i, n - i, lo, hi);
}
__attribute__((target("avx512f"))) void minmax_avx512(const int32_t *s, size_t n, int32_t *lo, int32_t *hi) {
  size_t i = 0;
  if (n >= 32) {
    __m512i a0 = _mm512_loadu_si512(s), a1 = _mm512_loadu_si512(s + 16), b0 = a0, b1 = a1;
    for (i = 32; i + 32 <= n; i += 32) {
      const __m512i x = _mm512_loadu_si512(s + i), y = _mm512_loadu_si512(s + i + 16);
      a0 = _mm512_min_epi32(a0, x), b0 = _mm512_max_epi32(b0, x);
      a1 = _mm512_min_epi32(a1, y), b1 = _mm512_max_epi32(b1, y);
    }
    const int32_t mn = _mm512_reduce_min_epi32(_mm512_min_epi32(a0, a1)), mx = _mm512_reduce_max_epi32(_mm512_max_epi32(b0, b1));
    *lo = mn < *lo ? mn : *lo;
    *hi = mx > *hi ? mx : *hi;
  }
  minmax_scalar(s + i, n - i, lo, hi);
}

struct Dispatch {
  copy_fn lines = copy_lines_sse2;
  gather_fn gather = gather_scalar;
  minmax_fn minmax = minmax_scalar;
  const char *isa = "sse2";
  Dispatch() {
    __builtin_cpu_init();
    const char *force = getenv("CFB_STAGE_ISA");  // sse2 | avx2 | avx512 (measurement / tests)
    const bool want512 = !force || !strcmp(force, "avx512"), want2 = !force || !strcmp(force, "avx2") || want512;
    if (want2 && __builtin_cpu_supports("avx2")) {
      lines = copy_lines_avx2;
      gather = gather_avx2;
      minmax = minmax_avx2;
      isa = "avx2";
    }
    if (want512 && __builtin_cpu_supports("avx512f")) {
      lines = copy_lines_avx512;
      minmax = minmax_avx512;
      isa = "avx512";
    }
  }
};
const Dispatch g_dispatch;

}  // namespace

extern "C" {

// Contiguous copy into a staging tile with non-temporal stores.  The caller issues one sfence before the
// tile is handed to the DMA engine (flush_tile).
void cfb_stage_copy(void *dst, const void *src, size_t bytes) {
  char *d = (char *)dst;
  const char *s = (const char *)src;
  if (bytes < 256) {
    memcpy(d, s, bytes);
    return;
  }
  const size_t head = (64 - ((uintptr_t)d & 63)) & 63;
  if (head) {
    memcpy(d, s, head);
    d += head;
    s += head;
    bytes -= head;
  }
  const size_t body = bytes & ~(size_t)63;
  g_dispatch.lines(d, s, body);
  if (bytes > body) memcpy(d + body, s + body, bytes - body);
}

// dst[i] = src[sel[i]], 4-byte elements.
void cfb_stage_gather32(void *dst, const void *src, const uint32_t *sel, size_t count) {
  g_dispatch.gather((uint32_t *)dst, (const uint32_t *)src, sel, count);
}

// *lo = min(*lo, src[0..count)), *hi = max(*hi, src[0..count)): the key range of a staged categorical vector.
void cfb_stage_minmax32(const int32_t *src, size_t count, int32_t *lo, int32_t *hi) { g_dispatch.minmax(src, count, lo, hi); }

const char *cfb_stage_isa(void) { return g_dispatch.isa; }

// Host-memory ceiling of the feed on THIS box: `threads` threads each copy their own `bytes_per_thread` source
// buffer (first-touched by the thread) into their own destination in 8 KB pieces, like a worker thread staging
// 2048-row vectors; returns the aggregate GB/s of payload (read bytes == written bytes).  nt = 1: the staging
// copy above; nt = 0: plain memcpy.  bench.py quotes it next to the end-to-end number.
double cfb_host_copy_ceiling(size_t bytes_per_thread, int threads, int nt, int reps) {
  if (threads < 1 || reps < 1 || bytes_per_thread < (1u << 16)) return 0.0;
  const size_t piece = 8192, n = bytes_per_thread / piece * piece;
  std::vector<char *> src(threads, nullptr), dst(threads, nullptr);
  std::vector<double> secs(threads, 0.0);
  auto work = [&](int t) {
    if (posix_memalign((void **)&src[t], 4096, n) || posix_memalign((void **)&dst[t], 4096, n)) return;
    memset(src[t], 1, n);
    memset(dst[t], 2, n);
    for (int r = -1; r < reps; r++) {  // r == -1: warm-up
      const auto t0 = std::chrono::steady_clock::now();
      for (size_t o = 0; o < n; o += piece) {
        if (nt) cfb_stage_copy(dst[t] + o, src[t] + o, piece);
        else memcpy(dst[t] + o, src[t] + o, piece);
      }
      _mm_sfence();
      if (r >= 0) secs[t] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
  };
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++) pool.emplace_back(work, t);
  for (auto &th : pool) th.join();
  double worst = 0.0;
  for (int t = 0; t < threads; t++) {
    worst = secs[t] > worst ? secs[t] : worst;
    free(src[t]);
    free(dst[t]);
  }
  return worst > 0.0 ? (double)n * reps * threads / worst / 1e9 : 0.0;
}

}  // extern "C"
