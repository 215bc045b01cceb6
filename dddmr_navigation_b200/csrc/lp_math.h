// lp_math.h — transcendental functions shared, bit for bit, by the sm_100a kernels and
// the CPU oracle's "shared" math mode.
//
// Why this exists: the reference's rollout calls glibc cosf/sinf inside a float
// recurrence (trajectory_generators/theories/dd_simple_trajectory_generator_theory.cpp:457-464)
// and glibc/Eigen/tf2 double sin/cos/asin/atan2 in the pose transform and the pure-pursuit
// critic (dd_simple…cpp:416, mpc_critics/models/pure_pursuit_model.cpp:96-101). CUDA's libm is
// not bit-identical to glibc, so both sides evaluate THESE functions instead:
//
//  * lpm_sinf / lpm_cosf restate the algorithm glibc >= 2.28 uses for sinf/cosf (double
//    range reduction + degree-7/8 polynomials). tools/check_sincosf_vs_glibc.cpp compares
//    them with the box's glibc over EVERY float |x| < 120: 0 mismatches against the non-FMA
//    build of glibc 2.39, 34 mismatches (all |x| >= 17.27) against its FMA ifunc variant.
//    Rollout headings are bounded by max_vel_theta*sim_time (a few rad), so on the path
//    lpm_sinf/lpm_cosf == glibc sinf/cosf exactly.
//  * lpm_sin / lpm_cos / lpm_asin / lpm_atan / lpm_atan2 are < 1 ulp double routines in
//    the classic Cody–Waite + minimax-polynomial form (coefficients as published in
//    FreeBSD msun / fdlibm). They are NOT bit-identical to glibc's; the oracle's "libm"
//    mode quantifies the difference (<= 1 ulp, tests/test_math.py).
//
// Every expression is written without fused multiply-add: the reference's x86-64 build has
// no FMA, the oracle is compiled with -ffp-contract=off and the device code with -fmad=false.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define LPM_HD __host__ __device__ __forceinline__
// the big double routines are called from a few sites only: one out-of-line copy keeps the kernels' code
// inside the instruction cache (same arithmetic either way)
#define LPM_HD_BIG static __host__ __device__ __noinline__
#else
#define LPM_HD static inline
#define LPM_HD_BIG static inline
#endif

namespace lpm {

LPM_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
LPM_HD uint64_t d2u(double d) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(d);
#else
  uint64_t u; memcpy(&u, &d, 8); return u;
#endif
}
LPM_HD double u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double d; memcpy(&d, &u, 8); return d;
#endif
}
LPM_HD double dabs(double x) { return u2d(d2u(x) & 0x7fffffffffffffffull); }
LPM_HD double dsqrt(double x) {
#if defined(__CUDA_ARCH__)
  return __dsqrt_rn(x);
#else
  return __builtin_sqrt(x);
#endif
}
LPM_HD float fsqrt(float x) {
#if defined(__CUDA_ARCH__)
  return __fsqrt_rn(x);
#else
  return __builtin_sqrtf(x);
#endif
}

// ----------------------------------------------------------------------------------------
// float sinf/cosf — glibc (sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, sincosf.h) restated.
// ----------------------------------------------------------------------------------------
struct sincosf_tab {
  double hpi_inv, hpi, c0, c1, c2, c3, c4, s1, s2, s3;
};

LPM_HD uint32_t abstop12(float x) { return (f2u(x) >> 20) & 0x7ff; }

// n even: sine polynomial; n odd: cosine polynomial. neg selects the negated-cosine table.
LPM_HD float sinf_poly(double x, double x2, bool neg, int n) {
  const double c0 = neg ? -0x1p0 : 0x1p0;
  const double c1 = neg ? 0x1.ffffffd0c621cp-2 : -0x1.ffffffd0c621cp-2;
  const double c2 = neg ? -0x1.55553e1068f19p-5 : 0x1.55553e1068f19p-5;
  const double c3 = neg ? 0x1.6c087e89a359dp-10 : -0x1.6c087e89a359dp-10;
  const double c4 = neg ? -0x1.99343027bf8c3p-16 : 0x1.99343027bf8c3p-16;
  const double s1 = -0x1.555545995a603p-3;
  const double s2 = 0x1.1107605230bc4p-7;
  const double s3 = -0x1.994eb3774cf24p-13;
  if ((n & 1) == 0) {
    double x3 = x * x2;
    double t1 = s2 + x2 * s3;
    double x7 = x3 * x2;
    double s = x + x3 * s1;
    return (float)(s + x7 * t1);
  } else {
    double x4 = x2 * x2;
    double t2 = c3 + x2 * c4;
    double t1 = c0 + x2 * c1;
    double x6 = x4 * x2;
    double c = t1 + x4 * c2;
    return (float)(c + x6 * t2);
  }
}

LPM_HD double reduce_fast(double x, int* np) {
  const double hpi_inv = 0x1.45F306DC9C883p+23;  // 2/pi * 2^24
  const double hpi = 0x1.921FB54442D18p0;        // pi/2
  double r = x * hpi_inv;
  int n = ((int32_t)r + 0x800000) >> 24;
  *np = n;
  return x - n * hpi;
}

LPM_HD_BIG double sin(double x);
LPM_HD_BIG double cos(double x);

// sinf and cosf of the same argument from ONE range reduction. glibc evaluates, for n = round(y / (pi/2)):
//   sinf: n even -> sine polynomial of (x*s, x^2), n odd -> cosine polynomial with the sign table (n & 2)
//   cosf: the same with n ^ 1
// where negating every coefficient of the cosine polynomial negates its value exactly. Its |y| < pi/4 shortcut is the
// n = 0 case of the general path (reduce_fast returns x unchanged), so one straight-line evaluation of both
// polynomials yields both results bit for bit; only the tiny-argument returns and the |y| >= 120 fallback remain.
LPM_HD void sincosf(float y, float* sn, float* cs) {
  const uint32_t top = abstop12(y);
  double x = y;
  if (top >= 0x42fu /* abstop12(120.0f) */) {
    // outside the planner's contract (the C ABI rejects such parameter sets); glibc switches to a 192-bit
    // reduction here, we round the double routines.
    *sn = (float)lpm::sin(x);
    *cs = (float)lpm::cos(x);
    return;
  }
  int n;
  const double xr = reduce_fast(x, &n);
  const double sg = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
  const double x2 = xr * xr;
  const float sp = sinf_poly(xr * sg, x2, false, 0);
  const float cp = sinf_poly(xr * sg, x2, false, 1);
  const float cpn = (n & 2) ? -cp : cp;
  float s = (n & 1) ? cpn : sp;
  float c = (n & 1) ? sp : cpn;
  if (top < 0x398u /* abstop12(2^-12) */) {
    s = y;
    c = 1.0f;
  }
  *sn = s;
  *cs = c;
}

LPM_HD float sinf(float y) {
  float s, c;
  sincosf(y, &s, &c);
  return s;
}

LPM_HD float cosf(float y) {
  float s, c;
  sincosf(y, &s, &c);
  return c;
}

// ----------------------------------------------------------------------------------------
// double sin/cos: Cody–Waite reduction by pi/2 (two 33-bit heads + tails, 118 bits of pi/2)
// followed by the msun/fdlibm kernels. Valid for |x| < 2^20*pi/2.
// ----------------------------------------------------------------------------------------
LPM_HD double ksin(double x, double y, int iy) {
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
               S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
               S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  double z = x * x;
  double v = z * x;
  double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
  if (iy == 0) return x + v * (S1 + z * r);
  return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}

LPM_HD double kcos(double x, double y) {
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
               C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
               C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  double z = x * x;
  double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
  double ax = dabs(x);
  if (ax < 0.3) return 1.0 - (0.5 * z - (z * r - x * y));
  double qx;
  if (ax > 0.78125) qx = 0.28125;
  else qx = u2d((d2u(ax) - 0x0020000000000000ull) & 0xffffffff00000000ull);  // ~|x|/4
  double hz = 0.5 * z - qx;
  double a = 1.0 - qx;
  return a - (hz - (z * r - x * y));
}

// returns quadrant n (mod 4 meaningful), y[0]+y[1] = x - n*pi/2
LPM_HD int rem_pio2(double x, double* y0, double* y1) {
  const double invpio2 = 6.36619772367581382433e-01;
  const double pio2_1 = 1.57079632673412561417e+00, pio2_1t = 6.07710050650619224932e-11;
  const double pio2_2 = 6.07710050630396597660e-11, pio2_2t = 2.02226624879595063154e-21;
  const double toint = 6755399441055744.0;  // 1.5 * 2^52
  double fn = (x * invpio2 + toint) - toint;
  double t = x - fn * pio2_1;  // exact: pio2_1 has 33 significant bits
  double w = fn * pio2_2;
  double r = t - w;
  w = fn * pio2_2t - ((t - r) - w);
  (void)pio2_1t;
  *y0 = r - w;
  *y1 = (r - *y0) - w;
  return (int)fn;
}

LPM_HD_BIG double sin(double x) {
  if (dabs(x) <= 0.78539816339744827900) {
    if (dabs(x) < 7.450580596923828125e-9 /* 2^-27 */) return x;
    return ksin(x, 0.0, 0);
  }
  double y0, y1;
  int n = rem_pio2(x, &y0, &y1);
  switch (n & 3) {
    case 0: return ksin(y0, y1, 1);
    case 1: return kcos(y0, y1);
    case 2: return -ksin(y0, y1, 1);
    default: return -kcos(y0, y1);
  }
}

LPM_HD_BIG double cos(double x) {
  if (dabs(x) <= 0.78539816339744827900) {
    if (dabs(x) < 7.450580596923828125e-9) return 1.0;
    return kcos(x, 0.0);
  }
  double y0, y1;
  int n = rem_pio2(x, &y0, &y1);
  switch (n & 3) {
    case 0: return kcos(y0, y1);
    case 1: return -ksin(y0, y1, 1);
    case 2: return -kcos(y0, y1);
    default: return ksin(y0, y1, 1);
  }
}

// sin and cos of the same argument from one reduction: exactly the values lpm::sin / lpm::cos return
LPM_HD_BIG void sincos(double x, double* sn, double* cs) {
  if (dabs(x) <= 0.78539816339744827900) {
    const bool tiny = dabs(x) < 7.450580596923828125e-9;
    *sn = tiny ? x : ksin(x, 0.0, 0);
    *cs = tiny ? 1.0 : kcos(x, 0.0);
    return;
  }
  double y0, y1;
  const int n = rem_pio2(x, &y0, &y1);
  const double s = ksin(y0, y1, 1), c = kcos(y0, y1);
  switch (n & 3) {
    case 0: *sn = s; *cs = c; break;
    case 1: *sn = c; *cs = -s; break;
    case 2: *sn = -s; *cs = -c; break;
    default: *sn = -c; *cs = s; break;
  }
}

// ----------------------------------------------------------------------------------------
// asin / atan / atan2 (msun e_asin.c, s_atan.c, e_atan2.c forms). Finite inputs only: the
// critic feeds matrix entries of a finite rotation.
// ----------------------------------------------------------------------------------------
LPM_HD_BIG double asin(double x) {
  const double pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17,
               pio4_hi = 7.85398163397448278999e-01;
  const double pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01,
               pS2 = 2.01212532134862925881e-01, pS3 = -4.00555345006794114027e-02,
               pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05;
  const double qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00,
               qS3 = -6.88283971605453293030e-01, qS4 = 7.70381505559019352791e-02;
  double ax = dabs(x);
  if (ax >= 1.0) {
    if (ax == 1.0) return x * pio2_hi + x * pio2_lo;
    return u2d(0x7ff8000000000000ull);
  }
  if (ax < 0.5) {
    if (ax < 7.450580596923828125e-9) return x;
    double t = x * x;
    double p = t * (pS0 + t * (pS1 + t * (pS2 + t * (pS3 + t * (pS4 + t * pS5)))));
    double q = 1.0 + t * (qS1 + t * (qS2 + t * (qS3 + t * qS4)));
    return x + x * (p / q);
  }
  double w = 1.0 - ax;
  double t = w * 0.5;
  double p = t * (pS0 + t * (pS1 + t * (pS2 + t * (pS3 + t * (pS4 + t * pS5)))));
  double q = 1.0 + t * (qS1 + t * (qS2 + t * (qS3 + t * qS4)));
  double s = dsqrt(t);
  if (ax >= 0.975) {
    w = p / q;
    t = pio2_hi - (2.0 * (s + s * w) - pio2_lo);
  } else {
    w = u2d(d2u(s) & 0xffffffff00000000ull);
    double c = (t - w * w) / (s + w);
    double r = p / q;
    p = 2.0 * s * r - (pio2_lo - 2.0 * c);
    q = pio4_hi - 2.0 * w;
    t = pio4_hi - (p - q);
  }
  return (x > 0.0) ? t : -t;
}

LPM_HD_BIG double atan(double x) {
  const double aT0 = 3.33333333333329318027e-01, aT1 = -1.99999999998764832476e-01,
               aT2 = 1.42857142725034663711e-01, aT3 = -1.11111104054623557880e-01,
               aT4 = 9.09088713343650656196e-02, aT5 = -7.69187620504482999495e-02,
               aT6 = 6.66107313738753120669e-02, aT7 = -5.83357013379057348645e-02,
               aT8 = 4.97687799461593236017e-02, aT9 = -3.65315727442169155270e-02,
               aT10 = 1.62858201153657823623e-02;
  bool neg = (d2u(x) >> 63) != 0;
  double ax = dabs(x);
  int id;
  double hi = 0.0, lo = 0.0;
  if (ax >= 7.378697629483820646e19 /* 2^66 */) {
    double z = 1.57079632679489655800e+00 + 6.12323399573676603587e-17;
    return neg ? -z : z;
  }
  if (ax < 0.4375) {
    if (ax < 1.862645149230957e-9 /* 2^-29 */) return x;
    id = -1;
  } else {
    x = ax;
    if (ax < 1.1875) {
      if (ax < 0.6875) {
        id = 0; x = (2.0 * x - 1.0) / (2.0 + x);
        hi = 4.63647609000806093515e-01; lo = 2.26987774529616870924e-17;
      } else {
        id = 1; x = (x - 1.0) / (x + 1.0);
        hi = 7.85398163397448278999e-01; lo = 3.06161699786838301793e-17;
      }
    } else {
      if (ax < 2.4375) {
        id = 2; x = (x - 1.5) / (1.0 + 1.5 * x);
        hi = 9.82793723247329054082e-01; lo = 1.39033110312309984516e-17;
      } else {
        id = 3; x = -1.0 / x;
        hi = 1.57079632679489655800e+00; lo = 6.12323399573676603587e-17;
      }
    }
  }
  double z = x * x;
  double w = z * z;
  double s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
  double s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
  if (id < 0) return x - x * (s1 + s2);
  z = hi - ((x * (s1 + s2) - lo) - x);
  return neg ? -z : z;
}

LPM_HD_BIG double atan2(double y, double x) {
  const double pi = 3.1415926535897931160E+00, pi_lo = 1.2246467991473531772E-16,
               pi_o_2 = 1.5707963267948965580E+00;
  const bool sx = (d2u(x) >> 63) != 0, sy = (d2u(y) >> 63) != 0;
  const int m = (sy ? 1 : 0) | (sx ? 2 : 0);
  if (x == 1.0) return lpm::atan(y);
  if (y == 0.0) {
    switch (m) {
      case 0: case 1: return y;
      case 2: return pi;
      default: return -pi;
    }
  }
  if (x == 0.0) return sy ? -pi_o_2 : pi_o_2;
  const int ey = (int)((d2u(y) >> 52) & 0x7ff), ex = (int)((d2u(x) >> 52) & 0x7ff);
  const int k = ey - ex;
  double z;
  if (k > 60) z = pi_o_2 + 0.5 * pi_lo;
  else if (sx && k < -60) z = 0.0;
  else z = lpm::atan(dabs(y / x));
  switch (m) {
    case 0: return z;
    case 1: return -z;
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
  }
}

// fmod for 0 <= x, y > 0, x < 2^52*y (exact, by repeated exact subtraction of scaled y is
// not needed on this path: the critic calls fmod(yaw + 3.1416, 3.1416) with yaw in [-pi, pi],
// so x < 2y and one Sterbenz-exact subtraction suffices). General inputs fall back to a loop.
LPM_HD double fmod_pos(double x, double y) {
  if (!(x >= 0.0) || !(y > 0.0)) return u2d(0x7ff8000000000000ull);
  if (x < y) return x;
  if (x < 2.0 * y) return x - y;  // exact (Sterbenz)
  // generic path: binary long division on the exponent gap (exact at every step)
  double r = x;
  while (r >= y) {
    double t = y;
    while (t + t <= r && t + t > t) t = t + t;
    r = r - t;  // exact: t <= r < 2t
  }
  return r;
}

}  // namespace lpm
