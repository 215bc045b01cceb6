"""Local_Planner::prunePlan (local_planner.cpp:374-445) and
perception_3d::PathBlockedStrategy::selfMark (path_blocked_strategy.cpp:56-100).

CPU part: the oracle restatement against hand-derived cases. GPU part: the device kernels against the oracle, and the
device-side prune plan feeding the cycle without leaving the GPU."""
import math

import numpy as np
import pytest

from dddmr_navigation_b200 import LocalPlanner, make_query, synth
from oracle import lporacle as O
from tests.helpers import assert_same_array, assert_trajectories_equal


def line_plan(n=100, step=0.1, x0=0.0, y=0.0, z=0.0):
    p = np.zeros((n, 7))
    p[:, 0] = x0 + step * np.arange(n)
    p[:, 1] = y
    p[:, 2] = z
    p[:, 6] = 1.0
    return p


def test_prune_walks_until_the_distances_go_negative_and_duplicates_the_nearest_pose():
    g = line_plan()
    info, poses, pcl = O.prune_plan(g, (2.04, 0.3, 0.0), forward_distance=1.0, backward_distance=0.5)
    assert info.status == 0 and info.nearest_index == 20
    # backward: 20, then 19..: 0.5 - 0.1*k < 0 first at k = 6 (0.1*5 is 0.5000000000000001 > 0.5 only after rounding; count it)
    back = []
    d, i, last = 0.5, 20, 20
    while i >= 0:
        back.append(i)
        if i < 20:
            d -= math.sqrt((g[last, 0] - g[i, 0]) ** 2)
        last = i
        if d < 0:
            break
        i -= 1
    fwd = []
    d, i = 1.0, 20
    while i < len(g):
        fwd.append(i)
        if i > 20:
            d -= math.sqrt((g[last, 0] - g[i, 0]) ** 2)
        last = i
        if d < 0:
            break
        i += 1
    assert info.n_backward == len(back) and info.n_prune == len(back) + len(fwd)
    assert np.array_equal(poses[:, 0], np.concatenate([g[back[::-1], 0], g[fwd, 0]]))   # poses: backward part reversed
    assert np.array_equal(pcl[:, 0], np.concatenate([g[back, 0], g[fwd, 0]]).astype(np.float32))  # pcl: as pushed
    assert np.all(pcl[:len(back), 3] == -1) and np.all(pcl[len(back):, 3] == 1)
    assert poses[len(back) - 1, 0] == poses[len(back), 0] == g[20, 0]  # the nearest pose appears twice, as upstream


def test_prune_intensity_zero_marks_the_first_pose_of_the_plan_and_short_or_distant_plans_bail_out():
    g = line_plan()
    info, poses, pcl = O.prune_plan(g, (0.02, 0.0, 0.0), 0.35, 1.0)
    assert info.nearest_index == 0 and info.n_backward == 1
    assert pcl[0, 3] == -1 and pcl[1, 3] == 0 and np.all(pcl[2:, 3] == 1)
    assert O.prune_plan(g[:2], (0, 0, 0), 1, 1)[0].status == 1
    far = O.prune_plan(g, (2.0, 1.0000001, 0.0), 1, 1)[0]
    assert far.status == 2 and far.n_prune == 0 or far.status == 2
    assert O.prune_plan(g, (2.0, 1.0, 0.0), 1, 1)[0].status == 0  # sqrtf(d2) > 1.0 is strict


def test_self_mark_counts_forward_points_with_a_strictly_closer_obstacle():
    cfg = synth.playground().config
    ora = O.OraclePlanner(cfg, O.MATH_SHARED, O.INDEX_BRUTE)
    cloud = np.zeros((8, 3), np.float32)
    cloud[:] = [50, 50, 50]
    cloud[0] = [1.0, 0.5, 0.0]        # exactly 0.5 from plan point x=1.0: NOT inside (strict <)
    cloud[1] = [2.0, 0.4999, 0.0]     # inside for x=2.0
    cloud[2] = [-1.0, 0.0, 0.0]       # on a backward point: ignored
    ora.set_cloud(cloud)
    pcl = np.array([[-1.0, 0, 0, -1], [0.0, 0, 0, -1], [0.0, 0, 0, 0], [1.0, 0, 0, 1], [2.0, 0, 0, 1], [3.0, 0, 0, 1]], np.float32)
    b = ora.path_blocked(pcl, 0.5)
    assert (b.n_blocked, b.n_checked, b.n_total, b.opinion) == (1, 4, 6, 1)
    assert b.ratio == float(np.float32(1) / np.float32(6)) * 100.0
    ora.set_cloud(cloud[:5])           # <= 5 points: the strategy reports 0 (path_blocked_strategy.cpp:62-64)
    b = ora.path_blocked(pcl, 0.5)
    assert b.ratio == 0.0 and b.opinion == 0


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["c1", "c2_doorway", "c3_ramp"])
def test_device_prune_and_self_mark_match_the_oracle_and_feed_the_cycle(case):
    if case == "c1":
        sc = synth.c1_ramp(n_points=50_000)
        pose, twist, plan = sc.pose, sc.twist, sc.plan
    elif case == "c2_doorway":
        sc = synth.c2_dense(n_points=300_000)
        pose, twist, plan = sc.pose, sc.twist, sc.plan
    else:
        sc = synth.c3_multilevel(n_points=600_000)
        pose, twist, plan = sc.extra_poses[1]
    # a long global plan: the scenario's prune plan continued straight ahead for another 6 m
    last, prev = plan[-1], plan[-2]
    d = (last[:3] - prev[:3])
    ext = np.repeat(last[None, :], 120, 0)
    ext[:, :3] += d[None, :] * np.arange(1, 121)[:, None]
    gplan = np.concatenate([plan, ext])
    gpu = LocalPlanner(sc.config)
    gpu.set_cloud(sc.cloud)
    gpu.set_global_plan(gplan)
    ora = O.OraclePlanner(sc.config, O.MATH_SHARED, O.INDEX_GRID)
    ora.set_cloud(sc.cloud)
    for fwd, bwd in [(3.0, 1.0), (1.0, 0.5), (0.0, 0.0), (50.0, 50.0)]:
        info = gpu.prune_plan(pose[:3], fwd, bwd)
        o_info, o_poses, o_pcl = O.prune_plan(gplan, pose[:3], fwd, bwd)
        assert info.as_dict() == o_info.as_dict(), (case, fwd, bwd)
        g_poses, g_pcl = gpu.read_prune_plan(info.n_prune)
        assert_same_array(g_poses, o_poses, "prune poses")
        assert_same_array(g_pcl, o_pcl, "pcl_prune_plan_")
        for radius in (0.2, 0.5, 1.3):
            b_g, b_o = gpu.path_blocked(radius), ora.path_blocked(o_pcl, radius)
            assert b_g.as_dict() == b_o.as_dict(), (case, fwd, bwd, radius)
        # the cycle runs on the device-side prune plan exactly as on the same plan handed over from the host
        q = make_query(pose, twist)
        r_dev = gpu.plan(q)
        t_dev = gpu.read_trajectories()
        ora.set_plan(o_poses)
        r_o = ora.plan(q)
        assert r_dev.as_dict() == r_o.as_dict(), (case, fwd, bwd)
        assert_trajectories_equal(t_dev, ora.read_trajectories())
    # robot off the plan: the prune plan is cleared and every trajectory is rejected by pure pursuit (-4)
    info = gpu.prune_plan([pose[0], pose[1] + 1.5, pose[2]], 3.0, 1.0)
    assert info.status == 2 and info.n_prune == 0
    assert gpu.plan(make_query(pose, twist)).best_id == -1
    # a host-side set_plan takes over again
    gpu.set_plan(plan)
    ora.set_plan(plan)
    assert gpu.plan(make_query(pose, twist)).as_dict() == ora.plan(make_query(pose, twist)).as_dict()
