// Exhaustive check of lpm::sinf / lpm::cosf (dddmr_navigation_b200/csrc/lp_math.h) against the glibc of
// the machine it runs on, over every float |x| < 120 (2.2e9 inputs, both signs), 8 threads, ~5 s.
//   g++ -O2 -ffp-contract=off -pthread -o /tmp/chk tools/check_sincosf_vs_glibc.cpp && /tmp/chk
//   GLIBC_TUNABLES=glibc.cpu.hwcaps=-FMA,-AVX2,-FMA4 /tmp/chk     # glibc's non-FMA ifunc variant
// Recorded on this image (glibc 2.39, x86-64): 0 mismatches vs the non-FMA variant; 12 (sinf) + 22 (cosf)
// vs the FMA variant, the smallest at |x| = 0x1.1475b6p+4 = 17.2787.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

#include "../dddmr_navigation_b200/csrc/lp_math.h"

int main() {
  uint32_t lim;
  const float f120 = 120.0f;
  memcpy(&lim, &f120, 4);
  std::atomic<uint64_t> bad_s{0}, bad_c{0}, tot{0};
  std::atomic<uint32_t> first_bad{0xffffffffu};
  std::vector<std::thread> th;
  const int NT = 8;
  for (int t = 0; t < NT; t++)
    th.emplace_back([&, t]() {
      uint64_t bs = 0, bc = 0, n = 0;
      for (uint64_t u = t; u < lim; u += NT)
        for (int sg = 0; sg < 2; sg++) {
          const uint32_t bits = (uint32_t)u | (sg ? 0x80000000u : 0);
          float y;
          memcpy(&y, &bits, 4);
          const float a = lpm::sinf(y), b = sinf(y), c = lpm::cosf(y), d = cosf(y);
          bool m = false;
          if (memcmp(&a, &b, 4)) { bs++; m = true; }
          if (memcmp(&c, &d, 4)) { bc++; m = true; }
          if (m) {
            uint32_t cur = first_bad.load();
            while ((uint32_t)u < cur && !first_bad.compare_exchange_weak(cur, (uint32_t)u)) {}
          }
          n++;
        }
      bad_s += bs; bad_c += bc; tot += n;
    });
  for (auto& x : th) x.join();
  printf("checked %lu floats |x|<120: sinf mismatches %lu, cosf mismatches %lu", (unsigned long)tot.load(),
         (unsigned long)bad_s.load(), (unsigned long)bad_c.load());
  if (first_bad.load() != 0xffffffffu) {
    const uint32_t fb = first_bad.load();
    float y;
    memcpy(&y, &fb, 4);
    printf(", smallest |x| with a mismatch = %a (%.6f)", y, y);
  }
  printf("\n");
  return 0;
}
