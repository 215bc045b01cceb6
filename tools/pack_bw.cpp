// pack_bw.cpp — how fast can T host threads pack a 32-byte-stride PointXYZI cloud into 12-byte xyz records in pinned memory,
// and how fast does the packed copy cross PCIe compared with the raw one? (feasibility probe for the host-side packing upload)
// build: nvcc -O3 -std=c++17 -Xcompiler -pthread -o tools/pack_bw tools/pack_bw.cpp
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include <immintrin.h>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static void pack_scalar(const float* src, float* dst, size_t i0, size_t i1) {
  for (size_t i = i0; i < i1; ++i) {
    const float* s = src + i * 8;
    float* d = dst + i * 3;
    d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
  }
}
template <bool kStream>
static void pack_sse(const float* src, float* dst, size_t i0, size_t i1) {
  size_t i = i0;
  for (; i + 4 <= i1; i += 4) {
    const float* s = src + i * 8;
    const __m128 a = _mm_loadu_ps(s), b = _mm_loadu_ps(s + 8), c = _mm_loadu_ps(s + 16), d = _mm_loadu_ps(s + 24);
    // o0 = a.x a.y a.z b.x ; o1 = b.y b.z c.x c.y ; o2 = c.z d.x d.y d.z
    const __m128 o0 = _mm_blend_ps(a, _mm_shuffle_ps(b, b, _MM_SHUFFLE(0, 0, 0, 0)), 0x8);
    const __m128 o1 = _mm_shuffle_ps(b, c, _MM_SHUFFLE(1, 0, 2, 1));
    const __m128 o2 = _mm_blend_ps(_mm_shuffle_ps(d, d, _MM_SHUFFLE(2, 1, 0, 0)), _mm_shuffle_ps(c, c, _MM_SHUFFLE(2, 2, 2, 2)), 0x1);
    float* o = dst + i * 3;
    if (kStream) { _mm_stream_ps(o, o0); _mm_stream_ps(o + 4, o1); _mm_stream_ps(o + 8, o2); }
    else { _mm_storeu_ps(o, o0); _mm_storeu_ps(o + 4, o1); _mm_storeu_ps(o + 8, o2); }
  }
  pack_scalar(src, dst, i, i1);
  if (kStream) _mm_sfence();
}
int main() {
  const size_t n = 2000000;
  float *raw, *packed; void *d_raw, *d_packed;
  cudaMallocHost(&raw, n * 32); cudaMallocHost(&packed, n * 12);
  cudaMalloc(&d_raw, n * 32); cudaMalloc(&d_packed, n * 12);
  for (size_t i = 0; i < n * 8; ++i) raw[i] = (float)i;
  std::vector<float> check(n * 3);
  pack_scalar(raw, check.data(), 0, n);
  typedef void (*fn_t)(const float*, float*, size_t, size_t);
  const fn_t fns[3] = {pack_scalar, pack_sse<false>, pack_sse<true>};
  const char* names[3] = {"scalar", "sse", "sse+stream"};
  for (int f = 0; f < 3; ++f)
    for (int T : {1, 4, 6, 8, 12, 16}) {
      double best = 1e9;
      for (int rep = 0; rep < 8; ++rep) {
        double t0 = now();
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t) {
          size_t a = (n * t / T) & ~(size_t)3, b = t == T - 1 ? n : (n * (t + 1) / T) & ~(size_t)3;
          th.emplace_back(fns[f], raw, packed, a, b);
        }
        for (auto& x : th) x.join();
        best = std::min(best, now() - t0);
      }
      printf("%-10s %2d threads (incl. spawn/join): %.3f ms  %s\n", names[f], T, best * 1e3, memcmp(packed, check.data(), n * 12) ? "MISMATCH" : "ok");
    }
  return 0;
}
