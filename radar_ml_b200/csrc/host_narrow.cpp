// see host_narrow.h
#include "host_narrow.h"

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include <immintrin.h>
#include <sched.h>

namespace rml_host {

int narrow_f32_to_u8_scalar(const float* src, uint8_t* dst, size_t n) {
  int bad = 0;
  for (size_t i = 0; i < n; ++i) {
    const float v = src[i];
    const int u = (v >= 0.f && v <= 255.f) ? static_cast<int>(v) : -1;   // NaN fails both comparisons
    bad |= (u < 0) | (static_cast<float>(u) != v);
    dst[i] = static_cast<uint8_t>(u);
  }
  return bad;
}

// 32 floats -> 32 bytes per step; the stores bypass the cache (the staging buffer is read next by DMA)
__attribute__((target("avx2"))) static int narrow_avx2(const float* s, uint8_t* d, size_t n) {
  __m256i bad = _mm256_setzero_si256();
  const __m256i hi = _mm256_set1_epi32(~0xFF);
  const __m256i order = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
  size_t i = 0;
  for (; i + 32 <= n; i += 32) {
    const __m256 a0 = _mm256_loadu_ps(s + i), a1 = _mm256_loadu_ps(s + i + 8);
    const __m256 a2 = _mm256_loadu_ps(s + i + 16), a3 = _mm256_loadu_ps(s + i + 24);
    const __m256i i0 = _mm256_cvtps_epi32(a0), i1 = _mm256_cvtps_epi32(a1);
    const __m256i i2 = _mm256_cvtps_epi32(a2), i3 = _mm256_cvtps_epi32(a3);
    // not an integer (or NaN): the round trip differs; out of range: bits above the low byte
    const __m256 n0 = _mm256_cmp_ps(_mm256_cvtepi32_ps(i0), a0, _CMP_NEQ_UQ);
    const __m256 n1 = _mm256_cmp_ps(_mm256_cvtepi32_ps(i1), a1, _CMP_NEQ_UQ);
    const __m256 n2 = _mm256_cmp_ps(_mm256_cvtepi32_ps(i2), a2, _CMP_NEQ_UQ);
    const __m256 n3 = _mm256_cmp_ps(_mm256_cvtepi32_ps(i3), a3, _CMP_NEQ_UQ);
    bad = _mm256_or_si256(bad, _mm256_castps_si256(_mm256_or_ps(_mm256_or_ps(n0, n1), _mm256_or_ps(n2, n3))));
    bad = _mm256_or_si256(bad, _mm256_and_si256(hi, _mm256_or_si256(_mm256_or_si256(i0, i1), _mm256_or_si256(i2, i3))));
    const __m256i p01 = _mm256_packus_epi32(i0, i1), p23 = _mm256_packus_epi32(i2, i3);
    const __m256i b = _mm256_permutevar8x32_epi32(_mm256_packus_epi16(p01, p23), order);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i), b);
  }
  _mm_sfence();
  int tail = (i < n) ? narrow_f32_to_u8_scalar(s + i, d + i, n - i) : 0;
  return tail | !_mm256_testz_si256(bad, bad);
}

static int narrow_range(const float* s, uint8_t* d, size_t n) {
  static const bool avx2 = __builtin_cpu_supports("avx2");
  return avx2 ? narrow_avx2(s, d, n) : narrow_f32_to_u8_scalar(s, d, n);
}

struct NarrowPool {
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  const float* src = nullptr;
  uint8_t* dst = nullptr;
  size_t n = 0;
  unsigned generation = 0;
  int pending = 0;
  std::atomic<int> bad{0};
  bool stop = false;
};

static void worker_main(NarrowPool* p, int me, int count) {
  unsigned seen = 0;
  for (;;) {
    const float* s;
    uint8_t* d;
    size_t n;
    {
      std::unique_lock<std::mutex> lk(p->mu);
      p->cv_work.wait(lk, [&] { return p->stop || p->generation != seen; });
      if (p->stop) return;
      seen = p->generation;
      s = p->src; d = p->dst; n = p->n;
    }
    // slices on 32-element boundaries: aligned non-temporal stores, no byte shared between threads
    const size_t blocks = (n + 31) / 32;
    const size_t b0 = blocks * me / count, b1 = blocks * (me + 1) / count;
    const size_t lo = b0 * 32, hi = b1 * 32 < n ? b1 * 32 : n;
    if (lo < hi && narrow_range(s + lo, d + lo, hi - lo)) p->bad.store(1, std::memory_order_relaxed);
    {
      std::lock_guard<std::mutex> lk(p->mu);
      if (--p->pending == 0) p->cv_done.notify_one();
    }
  }
}

NarrowPool* narrow_pool_create(int threads) {
  if (threads <= 0) {
    cpu_set_t set;
    CPU_ZERO(&set);
    threads = sched_getaffinity(0, sizeof(set), &set) == 0 ? CPU_COUNT(&set) : static_cast<int>(std::thread::hardware_concurrency());
    if (threads > 32) threads = 32;
  }
  if (threads < 1) threads = 1;
  NarrowPool* p = new NarrowPool();
  for (int t = 0; t < threads; ++t) p->workers.emplace_back(worker_main, p, t, threads);
  return p;
}

void narrow_pool_destroy(NarrowPool* p) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> lk(p->mu);
    p->stop = true;
  }
  p->cv_work.notify_all();
  for (auto& t : p->workers) t.join();
  delete p;
}

int narrow_pool_threads(const NarrowPool* p) { return p ? static_cast<int>(p->workers.size()) : 0; }

int narrow_f32_to_u8(NarrowPool* p, const float* src, uint8_t* dst, size_t n) {
  if (!p || n < 65536) return narrow_range(src, dst, n);
  std::unique_lock<std::mutex> lk(p->mu);
  p->src = src; p->dst = dst; p->n = n;
  p->bad.store(0, std::memory_order_relaxed);
  p->pending = static_cast<int>(p->workers.size());
  ++p->generation;
  p->cv_work.notify_all();
  p->cv_done.wait(lk, [&] { return p->pending == 0; });
  return p->bad.load(std::memory_order_relaxed);
}

}  // namespace rml_host
