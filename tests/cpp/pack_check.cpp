// pack_check.cpp — CPU check of the host side of the packing upload (dddmr_navigation_b200/csrc/lp_hostpack.h): the thread
// pool's packed rows and bounds against a scalar reference, over strides, sizes that are not multiples of 4, thread counts and
// clouds with NaN / inf points. No GPU needed. Exit code 0 = all cases pass.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <random>
#include <vector>

#include "../../dddmr_navigation_b200/csrc/lp_hostpack.h"

static int check(size_t n, size_t stride, int threads, unsigned seed, bool with_nonfinite) {
  std::mt19937 rng(seed);
  std::uniform_real_distribution<float> u(-50.f, 50.f);
  std::vector<char> raw(n * stride + 64, (char)0x7f);
  for (size_t i = 0; i < n; ++i) {
    float p[3] = {u(rng), u(rng), u(rng)};
    if (with_nonfinite && i % 97 == 5) p[i % 3] = std::numeric_limits<float>::quiet_NaN();
    if (with_nonfinite && i % 131 == 7) p[(i + 1) % 3] = (i & 1) ? INFINITY : -INFINITY;
    memcpy(raw.data() + i * stride, p, 12);
    const float pad = std::numeric_limits<float>::quiet_NaN();  // padding bytes must never reach the bounds
    if (stride >= 16) memcpy(raw.data() + i * stride + 12, &pad, 4);
  }
  // scalar reference
  float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f}, mx[3] = {-mn[0], -mn[0], -mn[0]};
  size_t n_finite = 0;
  std::vector<float> want(n * 3);
  for (size_t i = 0; i < n; ++i) {
    float p[3];
    memcpy(p, raw.data() + i * stride, 12);
    memcpy(&want[i * 3], p, 12);
    if (std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2])) {
      ++n_finite;
      for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], p[a]); mx[a] = std::max(mx[a], p[a]); }
    }
  }
  float* dst = nullptr;
  if (posix_memalign((void**)&dst, 64, n * 12 + 64)) return 1;
  lp::PackPool pool(threads);
  if (pool.threads() != threads) { std::printf("pool started %d of %d threads\n", pool.threads(), threads); return 1; }
  int bad = 0;
  for (int rep = 0; rep < 3 && !bad; ++rep) {  // the pool is reused across clouds
    memset(dst, 0xee, n * 12);
    pool.start(raw.data(), stride, dst, n);
    size_t covered = 0;
    for (int c = 0; c < lp::kPackChunks; ++c) {
      pool.wait_chunk(c);
      if (pool.bound(c) != covered || pool.bound(c + 1) < covered) bad = 1;
      covered = pool.bound(c + 1);
      // a chunk that has been waited for is complete
      if (memcmp(dst + pool.bound(c) * 3, want.data() + pool.bound(c) * 3, (pool.bound(c + 1) - pool.bound(c)) * 12)) bad = 2;
    }
    if (covered != n) bad = 3;
    const lp::HostBounds hb = pool.bounds();
    if (hb.n_finite != n_finite) bad = 4;
    if (n_finite)
      for (int a = 0; a < 3; ++a)
        if (hb.mn[a] != mn[a] || hb.mx[a] != mx[a]) bad = 5;
  }
  free(dst);
  if (bad) std::printf("FAIL n=%zu stride=%zu threads=%d nonfinite=%d: code %d\n", n, stride, threads, (int)with_nonfinite, bad);
  return bad;
}

int main() {
  int fails = 0, cases = 0;
  const size_t sizes[] = {0, 1, 3, 4, 5, 31, 32, 33, 1000, 4099, 65537, 300001};
  const size_t strides[] = {16, 20, 32, 48};
  const int threads[] = {1, 2, 3, 8};
  unsigned seed = 1;
  for (size_t n : sizes)
    for (size_t st : strides)
      for (int t : threads)
        for (int nf = 0; nf < 2; ++nf) {
          fails += check(n, st, t, seed++, nf != 0) ? 1 : 0;
          ++cases;
        }
  std::printf("%d cases, %d failed\n", cases, fails);
  return fails ? 1 : 0;
}
