"""lp_math.h (shared host/device transcendental functions) against glibc and against mpmath."""
import math
import struct

import mpmath
import numpy as np
import pytest

from oracle import lporacle as O

SH, LM = O.MATH_SHARED, O.MATH_LIBM


def _ulp(x: float) -> float:
    return math.ulp(x) if x != 0 else 5e-324


def test_sinf_cosf_match_glibc_bit_for_bit_on_the_rollout_range():
    """Headings on the path are bounded by max_vel_theta*sim_time (a few rad, 6.28 for rotate in place).
    Every 37th float in [0, 17] (both signs): shared == this box's glibc. (tools/check_sincosf_vs_glibc.cpp is the
    exhaustive version: 0 mismatches vs non-FMA glibc 2.39 over all |x| < 120, 34 vs the FMA ifunc, all |x| >= 17.27.)"""
    lib = O.load()
    hi = struct.unpack("<I", struct.pack("<f", 17.0))[0]
    assert lib.lporacle_sincosf_mismatches(0, hi, 37) == 0


def test_sinf_cosf_dense_window_exhaustive():
    lib = O.load()
    lo = struct.unpack("<I", struct.pack("<f", 0.5))[0]
    hi = struct.unpack("<I", struct.pack("<f", 0.5625))[0]
    assert lib.lporacle_sincosf_mismatches(lo, hi, 1) == 0  # every float in [0.5, 0.5625)


@pytest.mark.parametrize("fn,lo,hi", [("sin", -7.0, 7.0), ("cos", -7.0, 7.0), ("sin", -100.0, 100.0),
                                      ("cos", 1.5, 1.65), ("asin", -1.0, 1.0)])
def test_double_functions_within_one_ulp_of_exact(fn, lo, hi):
    lib = O.load()
    mpmath.mp.prec = 200
    rng = np.random.default_rng(7)
    f = getattr(lib, "lporacle_" + fn)
    exact = getattr(mpmath, fn)
    worst = 0.0
    for x in rng.uniform(lo, hi, 4000):
        got = f(SH, float(x))
        ref = exact(mpmath.mpf(float(x)))
        err = abs(mpmath.mpf(got) - ref) / mpmath.mpf(_ulp(float(ref)))
        worst = max(worst, float(err))
    assert worst < 1.0, f"{fn}: {worst} ulp"


def test_atan2_within_two_ulp_of_exact_all_quadrants():
    # atan(y/x): the quotient's rounding adds to atan's < 1 ulp; glibc's own atan2 is not correctly rounded either
    lib = O.load()
    mpmath.mp.prec = 200
    rng = np.random.default_rng(8)
    worst = 0.0
    for _ in range(6000):
        y, x = (float(v) for v in rng.uniform(-2, 2, 2) * 10.0 ** rng.integers(-6, 3))
        got = lib.lporacle_atan2(SH, y, x)
        ref = mpmath.atan2(mpmath.mpf(y), mpmath.mpf(x))
        worst = max(worst, float(abs(mpmath.mpf(got) - ref) / mpmath.mpf(_ulp(float(ref)))))
    assert worst < 2.0, worst
    assert lib.lporacle_atan2(SH, 0.0, -1.0) == math.pi and lib.lporacle_atan2(SH, -0.0, -1.0) == -math.pi
    assert lib.lporacle_atan2(SH, 1.0, 0.0) == math.pi / 2 and lib.lporacle_atan2(SH, 0.0, 1.0) == 0.0


def test_shared_vs_libm_double_functions_differ_by_at_most_one_ulp():
    lib = O.load()
    rng = np.random.default_rng(9)
    for fn in ("sin", "cos"):
        f = getattr(lib, "lporacle_" + fn)
        for x in rng.uniform(-7, 7, 20000):
            a, b = f(SH, float(x)), f(LM, float(x))
            assert abs(a - b) <= _ulp(b), (fn, x, a, b)


def test_fmod_is_exact_on_the_critic_range():
    lib = O.load()
    rng = np.random.default_rng(10)
    for yaw in list(rng.uniform(-math.pi, math.pi, 20000)) + [0.0, -0.0, math.pi, -math.pi, 1e-300, -1e-300]:
        v = float(yaw) + 3.1416
        assert lib.lporacle_fmod(SH, v, 3.1416) == math.fmod(v, 3.1416)
    for x, y in ((10.25, 3.0), (-10.25, 3.0), (1e10, 7.5), (5.0, 5.0)):
        assert lib.lporacle_fmod(SH, x, y) == math.fmod(x, y)
