"""Shared comparison helpers for the parity tests (GPU product vs CPU oracle)."""
from __future__ import annotations

import numpy as np

from dddmr_navigation_b200.config import make_query

INT_FIELDS = ("sample_index", "num_steps", "first_hit_pose")


def assert_same_array(a, b, name):
    """Value equality with NaNs in the same places (+0 == -0: signed zeros never reach a comparison)."""
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"{name}: shape {a.shape} vs {b.shape}"
    if a.dtype.kind == "f":
        nan_a, nan_b = np.isnan(a), np.isnan(b)
        assert np.array_equal(nan_a, nan_b), f"{name}: NaN pattern differs"
        bad = ~nan_a & (a != b)
    else:
        bad = a != b
    if bad.any():
        idx = np.argwhere(bad)[:5]
        raise AssertionError(f"{name}: {int(bad.sum())} of {a.size} entries differ; first at {idx.tolist()}: "
                             f"{a[tuple(idx[0])]!r} vs {b[tuple(idx[0])]!r}")


def assert_result_equal(r_gpu, r_ref, exact_cost=True):
    for f in ("best_id", "n_samples", "n_traj", "n_collided", "n_poses"):
        assert getattr(r_gpu, f) == getattr(r_ref, f), f"result.{f}: {getattr(r_gpu, f)} vs {getattr(r_ref, f)}"
    for f in ("best_cost", "xv", "yv", "thetav"):
        a, b = getattr(r_gpu, f), getattr(r_ref, f)
        if exact_cost:
            assert a == b, f"result.{f}: {a!r} vs {b!r}"
        else:
            assert abs(a - b) <= 1e-4 * max(1e-300, abs(b)), f"result.{f}: {a!r} vs {b!r}"


def assert_trajectories_equal(t_gpu: dict, t_ref: dict, exact=True, rtol=1e-4):
    for k in t_ref:
        if exact or k in INT_FIELDS:
            assert_same_array(t_gpu[k], t_ref[k], k)
        else:
            a, b = np.asarray(t_gpu[k], np.float64), np.asarray(t_ref[k], np.float64)
            assert np.array_equal(np.isnan(a), np.isnan(b)), f"{k}: NaN pattern differs"
            m = ~np.isnan(a)
            # per-critic float scores within 1e-4 relative (BASELINE.json north_star)
            assert np.all(np.abs(a[m] - b[m]) <= rtol * np.maximum(np.abs(b[m]), 1e-12)), f"{k}: beyond {rtol} relative"


def reference_argmin(cost: np.ndarray):
    """Local_Planner::getBestTrajectory (local_planner.cpp:447-480) on a cost array: min cost, ties -> LAST."""
    best, mc = -1, 9999999.0
    for i, c in enumerate(cost):
        if c >= 0 and c <= mc:
            best, mc = i, c
    return best


def run_pair(gpu, oracle, cloud, plan, pose, twist, max_speed=-1.0, heading_dev=0.0):
    q = make_query(pose, twist, max_speed, heading_dev)
    for p in (gpu, oracle):
        p.set_cloud(cloud)
        p.set_plan(plan)
    return gpu.plan(q), oracle.plan(q)
