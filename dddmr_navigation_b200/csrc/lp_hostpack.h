// lp_hostpack.h — host side of the packing upload of b200lp_set_cloud (no CUDA code in here except the device count the
// thread heuristic asks for): SSE packing of x,y,z with the bounds reduction, and the small thread pool that runs it
// chunk by chunk while b200lp.cu copies the chunks that are ready.
#pragma once
#include <cuda_runtime.h>
#include <immintrin.h>
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <ctime>
#include <mutex>
#include <thread>
#include <vector>

namespace lp {

// ---------------------------------------------------------------------------------------------------------------------
// Host-side packing upload. A pcl::PointXYZI cloud carries 12 useful bytes in every 32: a few host threads copy x,y,z
// of every point into a pinned staging buffer (streaming stores), chunk by chunk, while the chunks already packed cross
// PCIe — 24 MB instead of 64 MB for 2 M points, and the caller's buffer may be ordinary pageable memory (a plain
// cudaMemcpy from pageable memory runs at ~12 GB/s). Measured on the bench box (tools/pack_bw.cpp): 8 threads pack 2 M
// points in 0.40 ms, the packed copy takes 0.44 ms, the raw copy 1.16 ms (5.3 ms from pageable memory).
// ---------------------------------------------------------------------------------------------------------------------
#ifndef B200LP_PACK_CHUNKS
#define B200LP_PACK_CHUNKS 8
#endif
constexpr int kPackChunks = B200LP_PACK_CHUNKS;  // <= 16 (chunk_ev, CloudHeader)
#ifndef B200LP_PACK_GEOMETRIC
#define B200LP_PACK_GEOMETRIC 1  // 0: eight equal pieces (A/B builds)
#endif
#ifndef B200LP_PACK_MIN_BYTES
#define B200LP_PACK_MIN_BYTES (2u << 20)
#endif
constexpr size_t kPackMinBytes = B200LP_PACK_MIN_BYTES;  // smaller host clouds are copied as they are

struct alignas(64) HostBounds {  // one per packing thread (own cache line)
  float mn[4], mx[4];
  size_t n_finite;
  void reset() {
    for (int a = 0; a < 4; ++a) { mn[a] = 3.402823466e+38f; mx[a] = -3.402823466e+38f; }
    n_finite = 0;
  }
};

// Packs points [i0, i1) and folds the bounds of the finite ones into hb — the same reduction bounds_pack_kernel does on
// the device for clouds that are not packed on the host (min / max of floats: exact in any order).
inline void pack_xyz(const char* src, size_t stride, float* dst, size_t i0, size_t i1, HostBounds& hb) {
  size_t i = i0;
  __m128 mn = _mm_loadu_ps(hb.mn), mx = _mm_loadu_ps(hb.mx);
  size_t cnt = hb.n_finite;
  const __m128 zero = _mm_setzero_ps();
  auto fold = [&](__m128 p) {  // lanes x, y, z of one point; lane 3 is padding and is ignored everywhere
    if ((_mm_movemask_ps(_mm_cmpeq_ps(_mm_sub_ps(p, p), zero)) & 7) == 7) {
      mn = _mm_min_ps(mn, p);
      mx = _mm_max_ps(mx, p);
      ++cnt;
    }
  };
  if (((uintptr_t)(dst + i * 3) & 15) == 0) {
    for (; i + 4 <= i1; i += 4) {  // 4 points -> 3 x 16 bytes, written around the cache (the DMA engine is the only reader)
      const char* sp = src + i * stride;
      const __m128 a = _mm_loadu_ps((const float*)sp), b = _mm_loadu_ps((const float*)(sp + stride)),
                   c = _mm_loadu_ps((const float*)(sp + 2 * stride)), d = _mm_loadu_ps((const float*)(sp + 3 * stride));
      const __m128 t0 = _mm_shuffle_ps(a, b, _MM_SHUFFLE(0, 0, 2, 2));   // a.z a.z b.x b.x
      const __m128 o0 = _mm_shuffle_ps(a, t0, _MM_SHUFFLE(2, 0, 1, 0));  // a.x a.y a.z b.x
      const __m128 o1 = _mm_shuffle_ps(b, c, _MM_SHUFFLE(1, 0, 2, 1));   // b.y b.z c.x c.y
      const __m128 t2 = _mm_shuffle_ps(c, d, _MM_SHUFFLE(0, 0, 2, 2));   // c.z c.z d.x d.x
      const __m128 o2 = _mm_shuffle_ps(t2, d, _MM_SHUFFLE(2, 1, 2, 0));  // c.z d.x d.y d.z
      float* o = dst + i * 3;
      _mm_stream_ps(o, o0);
      _mm_stream_ps(o + 4, o1);
      _mm_stream_ps(o + 8, o2);
      // x - x is 0 for finite x and NaN otherwise: one test for the four points, the per-point path only when it fails
      const __m128 nf = _mm_add_ps(_mm_add_ps(_mm_sub_ps(a, a), _mm_sub_ps(b, b)), _mm_add_ps(_mm_sub_ps(c, c), _mm_sub_ps(d, d)));
      if ((_mm_movemask_ps(_mm_cmpeq_ps(nf, zero)) & 7) == 7) {
        mn = _mm_min_ps(_mm_min_ps(mn, a), _mm_min_ps(_mm_min_ps(b, c), d));
        mx = _mm_max_ps(_mm_max_ps(mx, a), _mm_max_ps(_mm_max_ps(b, c), d));
        cnt += 4;
      } else {
        fold(a); fold(b); fold(c); fold(d);
      }
    }
  }
  for (; i < i1; ++i) {
    const float* sp = (const float*)(src + i * stride);
    float* o = dst + i * 3;
    o[0] = sp[0];
    o[1] = sp[1];
    o[2] = sp[2];
    fold(_mm_set_ps(0.f, sp[2], sp[1], sp[0]));
  }
  _mm_sfence();
  _mm_storeu_ps(hb.mn, mn);
  _mm_storeu_ps(hb.mx, mx);
  hb.n_finite = cnt;
}

class PackPool {
 public:
  // Starts up to `threads` workers; fewer if the system refuses (threads() tells; 0 = unusable). Never throws.
  explicit PackPool(int threads) noexcept {
    for (int c = 0; c < kPackChunks; ++c) done_[c].store(0);
    try {
      bounds_.resize((size_t)threads);
      th_.reserve((size_t)threads);
      for (int w = 0; w < threads; ++w) th_.emplace_back([this, w] { run(w); });
    } catch (...) {
    }
    T_ = (int)th_.size();  // the workers read T_ only after the first start(), which happens after construction
  }
  ~PackPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      ++gen_;
      gen_live_.store(gen_, std::memory_order_release);
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  int threads() const { return T_; }
  // chunk c covers points [bound(c), bound(c+1)); every bound is a multiple of 4 points so that the 16-byte streaming
  // stores of neighbouring slices never share a destination line fragment
  // The pieces are not equal: the copy engine cannot start before the first piece is packed and the grid build cannot
  // finish before the last piece's histogram, so the first and the last piece are small (1/32 and 2/32 of the cloud) and
  // the middle ones carry the rest — packing is faster than the copy, so the link stays busy in between.
  size_t bound(int c) const {
    if (c >= kPackChunks) return n_;
#if B200LP_PACK_CHUNKS == 8 && B200LP_PACK_GEOMETRIC
    constexpr size_t cum[9] = {0, 1, 3, 7, 13, 19, 25, 30, 32};
    return (n_ * cum[c] / 32) & ~(size_t)3;
#else
    return (n_ * (size_t)c / kPackChunks) & ~(size_t)3;
#endif
  }
  void start(const char* src, size_t stride, float* dst, size_t n) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      src_ = src; stride_ = stride; dst_ = dst; n_ = n;
      for (int c = 0; c < kPackChunks; ++c) done_[c].store(0, std::memory_order_relaxed);
      const long long now = now_ns();
      hot_.store(last_start_ns_ != 0 && now - last_start_ns_ < 2 * spin_ns_, std::memory_order_relaxed);
      last_start_ns_ = now;
      ++gen_;
      gen_live_.store(gen_, std::memory_order_release);
    }
    cv_.notify_all();
  }
  void wait_chunk(int c) const {  // a chunk takes tens of microseconds: spin
    while (done_[c].load(std::memory_order_acquire) != T_) __builtin_ia32_pause();
  }
  // bounds of the finite points of the whole cloud; valid once the last chunk has been waited for
  HostBounds bounds() const {
    HostBounds r;
    r.reset();
    for (int w = 0; w < T_; ++w) {
      const HostBounds& b = bounds_[(size_t)w];
      for (int a = 0; a < 3; ++a) { r.mn[a] = std::min(r.mn[a], b.mn[a]); r.mx[a] = std::max(r.mx[a], b.mx[a]); }
      r.n_finite += b.n_finite;
    }
    return r;
  }

 private:
  void run(int w) {
    unsigned long long seen = 0;
    for (;;) {
      // Spin-then-park: a worker that has just packed a cloud polls for the next one for spin_ns_ before it sleeps on the
      // condition variable. Waking a parked thread costs 100-200 us on the bench hosts (B200LP_HOST_TRACE: the first
      // 2 MB piece was packed 107-197 us after start() with parked workers, ~20 us with polling ones) — a fifth of a
      // 2 M-point upload. Only while clouds really arrive back to back: the workers poll when the last two clouds came
      // within twice the window of each other, so a planner that uploads every 50 ms never burns a core waiting.
      // B200LP_PACK_SPIN_US sets the window (default 2000 us; 0 parks at once).
      if (spin_ns_ > 0 && hot_.load(std::memory_order_relaxed)) {
        const long long t0 = now_ns();
        while (gen_live_.load(std::memory_order_acquire) == seen && now_ns() - t0 < spin_ns_) __builtin_ia32_pause();
      }
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
      }
      bounds_[(size_t)w].reset();
      for (int c = 0; c < kPackChunks; ++c) {
        const size_t c0 = bound(c), c1 = bound(c + 1), len = c1 - c0;
        const size_t a = c0 + ((len * (size_t)w / T_) & ~(size_t)3), b = w == T_ - 1 ? c1 : c0 + ((len * (size_t)(w + 1) / T_) & ~(size_t)3);
        if (b > a) pack_xyz(src_, stride_, dst_, a, b, bounds_[(size_t)w]);
        done_[c].fetch_add(1, std::memory_order_release);
      }
    }
  }
  static long long now_ns() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (long long)ts.tv_sec * 1000000000ll + ts.tv_nsec;
  }
  static long long spin_window_ns() {
    const char* e = std::getenv("B200LP_PACK_SPIN_US");
    const long long us = e ? atoll(e) : 2000;
    return us < 0 ? 0 : (us > 1000000 ? 1000000 : us) * 1000;
  }
  const long long spin_ns_ = spin_window_ns();
  std::atomic<unsigned long long> gen_live_{0};  // copy of gen_ the spinning workers poll without the mutex
  std::atomic<bool> hot_{false};                 // the last two clouds arrived within 2 x spin_ns_ of each other
  long long last_start_ns_ = 0;
  int T_ = 0;
  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable cv_;
  unsigned long long gen_ = 0;
  bool stop_ = false;
  const char* src_ = nullptr;
  size_t stride_ = 0, n_ = 0;
  float* dst_ = nullptr;
  std::atomic<int> done_[kPackChunks];
  std::vector<HostBounds> bounds_;
};

// How many host threads pack a cloud. B200LP_PACK_THREADS decides when set (the embedding application knows how many
// planner processes share the host). Otherwise: 3/4 of the CPUs this process may run on, divided by the GPUs it can see (one
// planner process per GPU is the deployment this library is built for), at most 12 — and none when that leaves fewer than 4:
// a thread packs ~18 GB/s, so three of them are no faster than the raw copy, and with every GPU of a box uploading at once
// the host's memory bandwidth is the limit, which packing (read 32 B + write 12 B + DMA 12 B per point) only makes worse.
int pack_threads_wanted() {
  if (const char* e = std::getenv("B200LP_PACK_THREADS")) return std::max(0, std::min(64, atoi(e)));
  unsigned hw = std::thread::hardware_concurrency();
  cpu_set_t set;  // the CPUs this process may run on (a rank bound to its GPU's NUMA node sees only those)
  if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0) hw = (unsigned)CPU_COUNT(&set);
  int gpus = 1;
  if (cudaGetDeviceCount(&gpus) != cudaSuccess || gpus < 1) gpus = 1;
  const unsigned t = std::min(12u, hw / (unsigned)gpus * 3 / 4);
  return t >= 4 ? (int)t : 0;  // 0: plain copies of the caller's buffer
}


}  // namespace lp
