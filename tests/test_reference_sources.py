"""Pins the oracle to the REFERENCE'S OWN C++.

oracle/_ref/liblpref.so is built (oracle/Makefile, target `lpref`) from the reference's unmodified sources where they lie
under /root/reference — the three trajectory-generator theories, StackedGenerator, base_trajectory::Trajectory, the seven
critics, StackedScoringModel — against stand-ins (oracle/ref_shims/) for the third-party headers they include (Eigen,
PCL, tf2, rclcpp, pluginlib: not installed in this image) and the reference's vendored nanoflann as the kd-tree. These
tests run that code and the restatement (oracle/lp_oracle.cpp, libm math mode) on identical inputs and require IDENTICAL
bits for everything the reference exposes: the trajectory list, velocities, time_delta, number of poses, every pose,
cuboid vertex and AABB, every critic's return value, the accumulated cost and the selected trajectory.

What this pins: the reference's control flow, operand types (float vs double), evaluation order, early returns, loop
bounds, tie-breaks. What it cannot pin: the arithmetic INSIDE Eigen / PCL / FLANN / tf2, which both sides take from SURVEY.md
Appendix A.
"""
import copy
import math

import numpy as np
import pytest

from dddmr_navigation_b200 import PlannerConfig, make_query, synth
from dddmr_navigation_b200.config import (DD_ROTATE_INPLACE_DEFAULT, DD_SIMPLE_DEFAULT, OMNI_SIMPLE_CRITICS,
                                          OMNI_SIMPLE_DEFAULT, ROTATE_CRITICS)
from oracle import lporacle as O
from tests.helpers import assert_same_array

pytestmark = pytest.mark.skipif(not O.have_reference_sources(), reason="oracle/_ref/liblpref.so not built (needs /root/reference)")

FIELDS = ("vel", "num_steps", "time_delta", "cost", "critic_scores")


def _compare(cfg, cloud, plan, pose, twist, max_speed=-1.0, hdev=0.0, pose_trajs=6, tag=""):
    ref = O.ReferencePlanner(cfg)
    ora = O.OraclePlanner(cfg, O.MATH_LIBM, O.INDEX_BRUTE if len(cloud) < 30_000 else O.INDEX_GRID)
    q = make_query(pose, twist, max_speed, hdev)
    for p in (ref, ora):
        p.set_cloud(cloud)
        p.set_plan(plan)
    r_r, r_o = ref.plan(q), ora.plan(q)
    for f in ("best_id", "n_traj", "n_collided", "n_poses", "best_cost", "xv", "yv", "thetav"):
        assert getattr(r_r, f) == getattr(r_o, f), (tag, f, getattr(r_r, f), getattr(r_o, f))
    t_r, t_o = ref.read_trajectories(), ora.read_trajectories()
    for k in FIELDS:
        assert_same_array(t_o[k], t_r[k], f"{tag} {k}")
    n = r_o.n_traj
    ids = sorted(set(int(i) for i in np.linspace(0, max(n - 1, 0), pose_trajs))) if n else []
    for tid in ids:
        steps = int(t_o["num_steps"][tid])
        p_r, p_o = ref.read_poses(tid, steps), ora.read_poses(tid, steps)
        for k in ("pose", "pcl_pose", "cuboid", "aabb"):
            assert_same_array(p_o[k], p_r[k], f"{tag} traj {tid} {k}")
    return r_o


def test_playground_scenario():
    sc = synth.playground()
    r = _compare(sc.config, sc.cloud, sc.plan, sc.pose, sc.twist, pose_trajs=55, tag="playground")
    assert r.best_id == 51 and r.n_collided == 17


def test_c1_ramp_pitched_pose():
    sc = synth.c1_ramp(n_points=20_000)
    r = _compare(sc.config, sc.cloud, sc.plan, sc.pose, sc.twist, pose_trajs=12, tag="c1")
    assert r.n_collided > 0 and r.best_id >= 0


def test_big_cuboid_on_the_c2_scene():
    sc = synth.c2_dense(n_points=100_000, samples=(12.0, 14.0))
    r = _compare(sc.config, sc.cloud, sc.plan, sc.pose, sc.twist, pose_trajs=8, tag="c2")
    assert 0 < r.n_collided < r.n_traj


def test_omni_and_rotate_theories():
    cloud = synth.small_scene(9, n_points=4000)
    plan = np.array([[-0.3 + 0.1 * i, 0.0, 0.0, 0, 0, 0, 1.0] for i in range(30)])
    cfg = PlannerConfig(generator=copy.deepcopy(OMNI_SIMPLE_DEFAULT), critics=copy.deepcopy(OMNI_SIMPLE_CRITICS))
    r = _compare(cfg, cloud, plan, [0.1, -0.1, 0, *synth.quat_from_rpy(0, 0, 0.4)], [0.3, 0.1, 0.1], tag="omni")
    assert r.n_traj > 100
    cfg = PlannerConfig(generator=copy.deepcopy(DD_ROTATE_INPLACE_DEFAULT), critics=copy.deepcopy(ROTATE_CRITICS))
    for hd in (0.5, -0.5):
        r = _compare(cfg, synth.small_scene(13, n_points=2000, extent=1.5), plan, [0, 0, 0, 0, 0, 0, 1], [0, 0, 0], hdev=hd, tag="rotate")
        assert r.n_traj == 2 and r.n_poses == 2 * 126


def test_edge_cases_small_clouds_and_short_plans():
    cfg = PlannerConfig(generator=dict(copy.deepcopy(DD_SIMPLE_DEFAULT), linear_x_sample=4.0, angular_z_sample=5.0))
    four = synth.to_xyzi(np.array([[0.5, 0, 0.2]] * 4, np.float32))
    five = synth.to_xyzi(np.array([[0.5, 0, 0.2]] * 5, np.float32))
    line = np.array([[-0.3 + 0.1 * i, 0.0, 0.0, 0, 0, 0, 1.0] for i in range(30)])
    for ci, cloud in enumerate((four, five)):
        for pi, plan in enumerate((line, line[:2], np.zeros((0, 7)))):
            _compare(cfg, cloud, plan, [0, 0, 0, 0, 0, 0, 1], [0.3, 0, 0.0], tag=f"edge {ci}/{pi}")


@pytest.mark.parametrize("seed", range(48))
def test_randomised_scenes(seed):
    """The scene generator of tests/test_fuzz_gpu.py (random theory, critic stack, cuboid, tilted pose, plan length), minus
    the non-finite cloud points (the kd-tree stand-in, like FLANN, is not defined on NaN coordinates)."""
    from tests.test_fuzz_gpu import _scene
    cfg, cloud, plan, pose, twist, max_speed, hdev = _scene(seed)
    cloud = cloud[np.isfinite(cloud[:, :3]).all(axis=1)]
    _compare(cfg, cloud, plan, pose, twist, max_speed, hdev, pose_trajs=4, tag=f"fuzz {seed}")


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["playground", "c1", "c2_small", "omni", "fuzz3", "fuzz8", "fuzz17", "fuzz30"])
def test_gpu_against_the_reference_sources(case):
    """The sm_100a path against the reference's own C++ directly (north_star bar): identical trajectory list, pose
    counts, collision outcomes and selected id; per-critic scores and cost within 1e-4 relative (the reference calls
    glibc's double sin/cos/asin/atan2, the device lp_math.h's, <= 1 ulp apart)."""
    from dddmr_navigation_b200 import LocalPlanner
    hdev, ms = 0.0, -1.0
    if case == "playground":
        sc = synth.playground(); cfg, cloud, plan, pose, twist = sc.config, sc.cloud, sc.plan, sc.pose, sc.twist
    elif case == "c1":
        sc = synth.c1_ramp(n_points=50_000); cfg, cloud, plan, pose, twist = sc.config, sc.cloud, sc.plan, sc.pose, sc.twist
    elif case == "c2_small":
        sc = synth.c2_dense(n_points=200_000, samples=(16.0, 18.0)); cfg, cloud, plan, pose, twist = sc.config, sc.cloud, sc.plan, sc.pose, sc.twist
    elif case == "omni":
        cfg = PlannerConfig(generator=copy.deepcopy(OMNI_SIMPLE_DEFAULT), critics=copy.deepcopy(OMNI_SIMPLE_CRITICS))
        cloud = synth.small_scene(9, n_points=4000)
        plan = np.array([[-0.3 + 0.1 * i, 0.0, 0.0, 0, 0, 0, 1.0] for i in range(30)])
        pose, twist = [0.1, -0.1, 0, *synth.quat_from_rpy(0, 0, 0.4)], [0.3, 0.1, 0.1]
    else:
        from tests.test_fuzz_gpu import _scene
        cfg, cloud, plan, pose, twist, ms, hdev = _scene(int(case[4:]))
        cloud = cloud[np.isfinite(cloud[:, :3]).all(axis=1)]
    gpu, ref = LocalPlanner(cfg), O.ReferencePlanner(cfg)
    q = make_query(pose, twist, ms, hdev)
    for p in (gpu, ref):
        p.set_cloud(cloud)
        p.set_plan(plan)
    r_g, r_r = gpu.plan(q), ref.plan(q)
    assert (r_g.best_id, r_g.n_traj, r_g.n_collided, r_g.n_poses) == (r_r.best_id, r_r.n_traj, r_r.n_collided, r_r.n_poses)
    t_g, t_r = gpu.read_trajectories(), ref.read_trajectories()
    assert_same_array(t_g["num_steps"], t_r["num_steps"], "num_steps")
    assert_same_array(t_g["vel"], t_r["vel"], "vel")
    assert_same_array(t_g["time_delta"], t_r["time_delta"], "time_delta")
    for k in ("cost", "critic_scores"):
        a, b = t_g[k], t_r[k]
        assert np.array_equal(np.isnan(a), np.isnan(b)), k
        m = ~np.isnan(a)
        assert np.array_equal(a[m] < 0, b[m] < 0), k  # rejections (collision -1, pure pursuit -4) are discrete
        assert np.all(np.abs(a[m] - b[m]) <= 1e-4 * np.maximum(np.abs(b[m]), 1e-12)), k  # 1e-4 relative (BASELINE.json north_star)
    if r_r.best_id >= 0:
        assert abs(r_g.best_cost - r_r.best_cost) <= 1e-4 * abs(r_r.best_cost)
    n = r_r.n_traj
    for tid in sorted(set(int(i) for i in np.linspace(0, max(n - 1, 0), 4))) if n else []:
        steps = int(t_r["num_steps"][tid])
        p_g, p_r = gpu.read_poses(tid, steps), ref.read_poses(tid, steps)
        for k in ("pcl_pose", "cuboid", "aabb"):   # float outputs of double arithmetic: equal unless a 1-ulp libm difference straddles a float rounding
            assert np.allclose(p_g[k], p_r[k], rtol=1e-6, atol=1e-6), (tid, k)
        assert np.allclose(p_g["pose"], p_r["pose"], rtol=1e-12, atol=1e-12), tid


def test_path_blocked_strategy_opinion_matches_the_reference_plugin():
    """the restated selfMark against the reference's own PathBlockedStrategy plugin."""
    rng = np.random.default_rng(7)
    sc = synth.c1_ramp(n_points=20_000)
    ora = O.OraclePlanner(sc.config, O.MATH_SHARED, O.INDEX_BRUTE)
    seen = set()
    for trial in range(24):
        n = int(rng.choice([3, 4, 5, 6, 40, 400, 20_000]))
        cloud = sc.cloud[rng.choice(len(sc.cloud), size=n, replace=False)]
        info, poses, pcl = O.prune_plan(sc.plan, sc.pose[:3], float(rng.uniform(0.2, 3.0)), float(rng.uniform(0.0, 1.0)))
        radius = float(rng.choice([0.05, 0.3, 0.8, 2.0]))
        if trial % 2 == 0 and n > 5:  # drop one cloud point somewhere around a random (forward or backward) plan point
            k = int(rng.integers(0, len(pcl)))
            d = rng.normal(size=3)
            cloud = cloud.copy()
            cloud[0, :3] = pcl[k, :3] + d / np.linalg.norm(d) * rng.uniform(0.0, 2.0 * radius)
        ora.set_cloud(cloud)
        b = ora.path_blocked(pcl, radius)
        assert b.opinion == O.reference_path_blocked_opinion(cloud, pcl, radius), (trial, n, radius, b.as_dict())
        seen.add(b.opinion)
    assert seen == {0, 1}
    # <= 5 cloud points or an empty prune plan: PASS whatever the geometry (path_blocked_strategy.cpp:62-64)
    five = np.repeat(np.array([[pcl[-1, 0], pcl[-1, 1], pcl[-1, 2]]], np.float32), 5, 0)
    assert O.reference_path_blocked_opinion(synth.to_xyzi(five), pcl, 1.0) == 0
    ora.set_cloud(synth.to_xyzi(five))
    assert ora.path_blocked(pcl, 1.0).opinion == 0
    six = np.repeat(five[:1], 6, 0)
    assert O.reference_path_blocked_opinion(synth.to_xyzi(six), pcl, 1.0) == 1
    ora.set_cloud(synth.to_xyzi(six))
    assert ora.path_blocked(pcl, 1.0).opinion == 1
