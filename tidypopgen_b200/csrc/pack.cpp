// Host side of the .bk streamer: a persistent worker pool and the 2-bit packer.
//
// The three pairwise kernels of the reference read the RAW FBM byte and treat everything
// above 2 as missing (src/snp_king.cpp:45-58, src/snp_ibs.cpp:45-56, src/snp_as.cpp:44-53),
// so a genotype carries two bits of information for this path: 0, 1, 2 or "missing".
// Packing four genotypes per byte while the chunk is staged into pinned memory (the pass
// over the bytes that staging makes anyway) cuts the PCIe traffic — the bound of every
// host-sourced job — by 4x; decode_planes_pk2 expands the codes on the GPU.
//
// Packed chunk layout ("pk2"): column (locus) k of the chunk occupies pk_stride bytes at
// k * pk_stride, pk_stride = round_up(ceil(span / 4), 16); byte b holds the individuals
// 4b .. 4b+3 of the staged row window, individual 4b+j in bits [2j, 2j+2); code 3 = missing
// (also used for the padding individuals of the last byte and the padding bytes).
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include <immintrin.h>
#include "hostpool.h"
#include "config.h"

namespace tpg {

static double now_ms_() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// ------------------------------------------------------------------ worker pool
struct HostPool::Impl {
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  std::vector<std::thread> workers;
  const std::function<void(int64_t, int64_t)>* body = nullptr;
  std::function<void(int64_t, int64_t)> owned;  // body of an asynchronous job (begin / end)
  double t_begin = 0, t_done = 0;
  int64_t n = 0;
  int parts = 0;
  int next = 0, done = 0;
  uint64_t epoch = 0;
  bool stop = false;

  void worker() {
    uint64_t seen = 0;
    std::unique_lock<std::mutex> lk(mu);
    for (;;) {
      cv_work.wait(lk, [&] { return stop || (epoch != seen && next < parts); });
      if (stop) return;
      while (next < parts) {
        const int t = next++;
        const auto* b = body;
        const int64_t a = n * t / parts, e = n * (t + 1) / parts;
        lk.unlock();
        (*b)(a, e);
        lk.lock();
        if (++done == parts) { t_done = now_ms_(); cv_done.notify_all(); }
      }
      seen = epoch;
    }
  }
};

HostPool::HostPool() : impl_(new Impl) {}

HostPool& HostPool::instance() {
  static HostPool* p = new HostPool();  // leaked on purpose: workers may outlive static destructors
  return *p;
}

int HostPool::max_threads() {
  unsigned hw = std::thread::hardware_concurrency();
  if (hw == 0) hw = 4;
  { const int v = cfg().host_threads; if (v >= 1) return v; }
  // one process per GPU under torchrun: the ranks of a node share its cores
  if (const char* e = getenv("LOCAL_WORLD_SIZE")) {
    const int w = atoi(e);
    if (w > 1) hw = std::max(2u, hw / unsigned(w));
  }
  return int(hw > 32 ? 32 : hw);
}

void HostPool::run(int64_t n, int64_t grain, const std::function<void(int64_t, int64_t)>& body) {
  if (n <= 0) return;
  int parts = int(std::min<int64_t>(max_threads(), std::max<int64_t>(1, n / std::max<int64_t>(grain, 1))));
  if (parts <= 1) { body(0, n); return; }
  Impl& s = *impl_;
  std::unique_lock<std::mutex> lk(s.mu);
  // one job at a time (callers are the per-device streamer threads)
  s.cv_done.wait(lk, [&] { return s.body == nullptr; });
  while (int(s.workers.size()) < parts - 1) s.workers.emplace_back([&s] { s.worker(); });
  s.body = &body;
  s.n = n;
  s.parts = parts;
  s.next = 0;
  s.done = 0;
  ++s.epoch;
  s.cv_work.notify_all();
  // the caller works too
  while (s.next < s.parts) {
    const int t = s.next++;
    const int64_t a = n * t / parts, e = n * (t + 1) / parts;
    lk.unlock();
    body(a, e);
    lk.lock();
    ++s.done;
  }
  s.cv_done.wait(lk, [&] { return s.done == s.parts; });
  s.body = nullptr;
  s.cv_done.notify_all();
}

void HostPool::begin(int64_t n, int64_t grain, std::function<void(int64_t, int64_t)> body) {
  Impl& s = *impl_;
  std::unique_lock<std::mutex> lk(s.mu);
  s.cv_done.wait(lk, [&] { return s.body == nullptr; });
  int parts = int(std::min<int64_t>(std::max(1, max_threads() - 1), std::max<int64_t>(1, n / std::max<int64_t>(grain, 1))));
  if (n <= 0) parts = 0;
  while (int(s.workers.size()) < std::max(parts, 1)) s.workers.emplace_back([&s] { s.worker(); });
  s.owned = std::move(body);
  s.body = &s.owned;
  s.n = n;
  s.parts = parts;
  s.next = 0;
  s.done = 0;
  s.t_begin = s.t_done = now_ms_();
  ++s.epoch;
  s.cv_work.notify_all();
}

double HostPool::end() {
  Impl& s = *impl_;
  std::unique_lock<std::mutex> lk(s.mu);
  s.cv_done.wait(lk, [&] { return s.done == s.parts; });
  const double ms = s.t_done - s.t_begin;
  s.body = nullptr;
  s.owned = nullptr;
  s.cv_done.notify_all();
  return ms;
}

// ------------------------------------------------------------------ 2-bit packer
static inline uint8_t pack4_scalar(const uint8_t* g, int64_t nv) {
  uint8_t b = 0xFF;  // all four = missing
  for (int64_t j = 0; j < nv; ++j) {
    const uint8_t c = g[j] > 2 ? 3 : g[j];
    b = uint8_t((b & ~(3u << (2 * j))) | (c << (2 * j)));
  }
  return b;
}

static void pack_col_scalar(const uint8_t* src, int64_t span, uint8_t* dst) {
  const int64_t full = span / 4;
  for (int64_t b = 0; b < full; ++b) dst[b] = pack4_scalar(src + 4 * b, 4);
  if (span % 4) dst[full] = pack4_scalar(src + 4 * full, span % 4);
}

// 32 raw bytes -> 8 packed bytes (as two dwords: individuals 0-15, 16-31)
__attribute__((target("avx2"))) static inline void pack32_avx2(__m256i x, uint32_t* lo, uint32_t* hi) {
  const __m256i three = _mm256_set1_epi8(3);
  const __m256i k1 = _mm256_set1_epi16(0x0401);
  const __m256i k2 = _mm256_set1_epi32(0x00100001);
  x = _mm256_min_epu8(x, three);
  const __m256i z = _mm256_madd_epi16(_mm256_maddubs_epi16(x, k1), k2);
  const __m256i w = _mm256_packus_epi32(z, z);
  const __m256i by = _mm256_packus_epi16(w, w);
  *lo = uint32_t(_mm256_extract_epi32(by, 0));
  *hi = uint32_t(_mm256_extract_epi32(by, 4));
}

__attribute__((target("avx2"))) static void pack_col_avx2(const uint8_t* src, int64_t span, uint8_t* dst) {
  const __m256i three = _mm256_set1_epi8(3);
  const __m256i k1 = _mm256_set1_epi16(0x0401);      // bytes (1, 4): b0 + 4 b1
  const __m256i k2 = _mm256_set1_epi32(0x00100001);  // words (1, 16): w0 + 16 w1
  const __m256i perm = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
  int64_t i = 0;
  for (; i + 128 <= span; i += 128) {
    __m256i z[4];
    for (int v = 0; v < 4; ++v) {
      __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32 * v));
      x = _mm256_min_epu8(x, three);
      z[v] = _mm256_madd_epi16(_mm256_maddubs_epi16(x, k1), k2);  // 8 dwords, one packed byte each
    }
    const __m256i p01 = _mm256_packus_epi32(z[0], z[1]);
    const __m256i p23 = _mm256_packus_epi32(z[2], z[3]);
    const __m256i by = _mm256_packus_epi16(p01, p23);  // lanes: [z0.lo z1.lo z2.lo z3.lo | z0.hi .. z3.hi]
    _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + i / 4), _mm256_permutevar8x32_epi32(by, perm));
  }
  for (; i + 32 <= span; i += 32) {
    uint32_t lo, hi;
    pack32_avx2(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i)), &lo, &hi);
    memcpy(dst + i / 4, &lo, 4);
    memcpy(dst + i / 4 + 4, &hi, 4);
  }
  if (i < span) {  // < 32 left: pad with "missing" in a local buffer
    alignas(32) uint8_t buf[32];
    memset(buf, 3, sizeof buf);
    memcpy(buf, src + i, size_t(span - i));
    uint32_t w[2];
    pack32_avx2(_mm256_load_si256(reinterpret_cast<const __m256i*>(buf)), &w[0], &w[1]);
    memcpy(dst + i / 4, w, size_t((span - i + 3) / 4));
  }
}

// memcpy for result pieces leaving the pinned bounce buffers: non-temporal stores.  The destination (a
// 20 GB R matrix at n = 50,000) is written once and not read again by us, so fetching its cache lines for
// ownership — what cached stores of a plain memcpy of this size per thread do — is a third of the DRAM
// traffic of the copy for nothing.
__attribute__((target("avx2"))) static void copy_stream_avx2(uint8_t* dst, const uint8_t* src, size_t n) {
  size_t head = (32 - (reinterpret_cast<uintptr_t>(dst) & 31)) & 31;
  if (head > n) head = n;
  memcpy(dst, src, head);
  dst += head; src += head; n -= head;
  size_t i = 0;
  for (; i + 128 <= n; i += 128) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32));
    const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 64));
    const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 96));
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), a);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 32), b);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 64), c);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 96), d);
  }
  _mm_sfence();
  memcpy(dst + i, src + i, n - i);
}

static bool have_avx2() {
  static const bool v = __builtin_cpu_supports("avx2");
  return v;
}

void copy_streaming(void* dst, const void* src, size_t n) {
  if (n >= 4096 && have_avx2()) copy_stream_avx2(static_cast<uint8_t*>(dst), static_cast<const uint8_t*>(src), n);
  else memcpy(dst, src, n);
}

int64_t pk2_stride(int64_t span) { return ((span + 3) / 4 + 15) / 16 * 16; }

void pk2_pack_column(const uint8_t* src, int64_t span, uint8_t* dst, int64_t stride) {
  if (have_avx2()) pack_col_avx2(src, span, dst);
  else pack_col_scalar(src, span, dst);
  const int64_t used = (span + 3) / 4;
  if (stride > used) memset(dst + used, 0xFF, size_t(stride - used));
}

}  // namespace tpg

extern "C" {

/* Packs ncol columns (column c at src + c*src_stride, span bytes each) into the pk2 layout at
 * dst (tpg_pk2_stride(span) bytes per column) on the library's host worker pool. */
int64_t tpg_pk2_stride(int64_t span) { return tpg::pk2_stride(span); }

void tpg_pk2_pack_host(const uint8_t* src, int64_t src_stride, int64_t span, int64_t ncol, uint8_t* dst) {
  const int64_t stride = tpg::pk2_stride(span);
  tpg::HostPool::instance().run(ncol, std::max<int64_t>(1, (int64_t(1) << 20) / std::max<int64_t>(span, 1)),
                                [=](int64_t a, int64_t e) {
                                  for (int64_t c = a; c < e; ++c)
                                    tpg::pk2_pack_column(src + c * src_stride, span, dst + c * stride, stride);
                                });
}

}  // extern "C"
