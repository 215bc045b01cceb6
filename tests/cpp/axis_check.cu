// axis_check.cu — CPU check of the closed-form velocity axes (dddmr_navigation_b200/csrc/lp_kernels.cuh: AxisPlan, axis_value,
// velocity_iterator_dev): for many windows and sample counts the VelocityIterator chain (what the reference runs,
// trajectory_generators/velocity_iterator.h:44-69) is compared entry by entry, bit for bit, with the closed form a
// single-robot launch hands prep_kernel. Counts how many entries need an exception (plan_samples lists up to 8 per launch and
// otherwise lets the kernel run the chains). Host code only; built with nvcc because the header is CUDA.
//   nvcc -std=c++17 -O2 -Xcompiler -ffp-contract=off -o axis_check tests/cpp/axis_check.cu && ./axis_check
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>

#include "../../dddmr_navigation_b200/csrc/lp_hostpack.h"
#include "../../dddmr_navigation_b200/csrc/lp_kernels.cuh"

using namespace lp;

static int plan_axis(float mn, float mx, int want, const float* chain, int n, AxisPlan* A) {  // as plan_samples does
  A->mn = (double)mn;
  A->step = 0.0;
  A->n_out = n;
  A->zero_at = -1;
  A->pad = 0;
  A->last = chain[n - 1];
  const int w = want < 2 ? 2 : want;
  if (mn != mx) A->step = ((double)mx - (double)mn) / (double)(w - 1);
  if (n == w + 1)
    for (int i = 1; i < n - 1; ++i)
      if (chain[i] == 0.0f && chain[i - 1] < 0.0f) { A->zero_at = i; break; }
  int exceptions = 0;
  for (int i = 0; i < n; ++i) {
    const float v = axis_value(*A, i);
    if (memcmp(&v, &chain[i], sizeof(float)) != 0) ++exceptions;
  }
  return exceptions;
}

int main() {
  std::mt19937_64 rng(20261019);
  std::uniform_real_distribution<double> u(-1.0, 1.0);
  static float chain[kMaxAxis];
  long long windows = 0, entries = 0, exc_entries = 0, exc_windows = 0, over_limit = 0, failed = 0;
  int worst = 0;
  const int counts[] = {1, 2, 3, 5, 20, 21, 25, 64, 128, 129, 361, 362, 500, 1000, 2045};
  for (int rep = 0; rep < 4000; ++rep)
    for (int want : counts) {
      // windows like the theories build them: limits and twist +- acceleration x period, narrowed to float
      double a = u(rng) * 2.0, b = u(rng) * 2.0;
      if (rep % 7 == 0) a = -b;          // symmetric: the chain passes (close to) zero
      if (rep % 11 == 0) b = a;          // degenerate window
      if (rep % 13 == 0) { a = 0.0; }    // starts at zero
      const float mn = (float)(a < b ? a : b), mx = (float)(a < b ? b : a);
      const int n = velocity_iterator_dev((double)mn, (double)mx, want, chain);
      AxisPlan A;
      const int e = plan_axis(mn, mx, want, chain, n, &A);
      ++windows;
      entries += n;
      exc_entries += e;
      exc_windows += e ? 1 : 0;
      over_limit += e > kAxisExceptions ? 1 : 0;
      if (e > worst) worst = e;

      // the layout claims of the closed form: first entry is the minimum, last the maximum, one inserted zero at most
      if (n < 1 || chain[n - 1] != mx || (mn != mx && chain[0] != mn)) {
        if (++failed <= 5) printf("  layout: mn %.9g mx %.9g want %d n %d first %.9g last %.9g\n", mn, mx, want, n, chain[0], chain[n - 1]);
      }
      // closed form + exceptions == chain, whatever the window: rebuild the axis the way prep_kernel does
      if (e <= kAxisExceptions) {
        int n_exc = 0, exc_at[kAxisExceptions];
        float exc_val[kAxisExceptions];
        for (int i = 0; i < n; ++i) {
          const float v = axis_value(A, i);
          if (memcmp(&v, &chain[i], sizeof(float)) != 0) { exc_at[n_exc] = i; exc_val[n_exc++] = chain[i]; }
        }
        static float rebuilt[kMaxAxis];
        for (int i = 0; i < n; ++i) rebuilt[i] = axis_value(A, i);
        for (int k = 0; k < n_exc; ++k) rebuilt[exc_at[k]] = exc_val[k];
        if (memcmp(rebuilt, chain, (size_t)n * sizeof(float)) != 0) ++failed;
      }
      // (a window that ends at 0 gets its zero twice — the inserted one and the maximum —, as upstream)
      if (A.zero_at >= 0 && !(chain[A.zero_at] == 0.0f && chain[A.zero_at - 1] < 0.0f && chain[A.zero_at + 1] >= 0.0f)) {
        if (++failed <= 10) printf("  zero: mn %.9g mx %.9g want %d n %d zero_at %d: %.9g %.9g %.9g\n", mn, mx, want, n, A.zero_at, chain[A.zero_at - 1], chain[A.zero_at], chain[A.zero_at + 1]);
      }
    }
  printf("%lld windows, %lld entries: %lld entries in %lld windows differ from the closed form (most in one window: %d), "
         "%lld windows beyond the %d exceptions a launch carries\n", windows, entries, exc_entries, exc_windows, worst, over_limit,
         kAxisExceptions);
  printf("%lld failed\n", failed);
  return failed ? 1 : 0;
}
