"""The C++ host layer (dddmr_navigation_b200/host): the reference's plugin interfaces over the C ABI.

tests/cpp/plugin_cycle drives a cycle exactly like Local_Planner::computeVelocityCommand (local_planner.cpp:482-621):
Trajectory_Generators_ROS / MPC_Critics_ROS load the plugins named in a ROS params YAML, the generator hands out
base_trajectory::Trajectory objects one at a time, StackedScoringModel sums the critics with its negative early-out
and getBestTrajectory keeps the last minimum. The GPU tests compare everything that caller sees with the CPU oracle.
"""
from __future__ import annotations

import math
import os
import struct
import subprocess

import numpy as np
import pytest

from dddmr_navigation_b200 import PlannerConfig, make_query, synth
from dddmr_navigation_b200 import config as cfgmod

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "plugin_cycle")


@pytest.fixture(scope="module")
def driver():
    import __graft_entry__ as g
    g.build_host()
    assert os.path.exists(BIN)
    return BIN


def _write_case(tmp, cfg, gen, cloud, plan, pose, twist, max_speed=-1.0, heading_dev=0.0):
    yaml = os.path.join(tmp, "params.yaml")
    with open(yaml, "w") as f:
        f.write(cfg.to_ros_yaml(gen))
    cloud = np.ascontiguousarray(cloud, np.float32)
    if cloud.shape[1] != 8:  # PointXYZI layout: x y z pad intensity pad pad pad
        c8 = np.zeros((cloud.shape[0], 8), np.float32)
        c8[:, :3] = cloud[:, :3]
        c8[:, 3] = 1.0
        cloud = c8
    plan = np.ascontiguousarray(plan, np.float64).reshape(-1, 7)
    sc = os.path.join(tmp, "scenario.bin")
    with open(sc, "wb") as f:
        f.write(struct.pack("<qq", cloud.shape[0], plan.shape[0]))
        f.write(cloud.tobytes())
        f.write(plan.tobytes())
        f.write(np.asarray(list(pose) + list(twist) + [max_speed, heading_dev], np.float64).tobytes())
    return yaml, sc


def _run(driver, yaml, sc, prefix, gen, mode):
    p = subprocess.run([driver, yaml, sc, prefix, gen, mode], capture_output=True, text=True, timeout=600)
    return p


def _load(prefix):
    s = {}
    for line in open(prefix + ".summary.txt"):
        k, v = line.strip().split("=")
        s[k] = float(v)
    traj = np.fromfile(prefix + ".traj.f64", np.float64).reshape(-1, 6)
    return s, traj, {"pose": np.fromfile(prefix + ".pose.f64", np.float64).reshape(-1, 7),
                     "pcl_pose": np.fromfile(prefix + ".pcl.f32", np.float32).reshape(-1, 3),
                     "cuboid": np.fromfile(prefix + ".cuboid.f32", np.float32).reshape(-1, 8, 3),
                     "aabb": np.fromfile(prefix + ".aabb.f32", np.float32).reshape(-1, 6)}


def test_host_layer_builds_and_fails_loudly_without_a_device(driver, tmp_path):
    """No GPU here: the plugins load from the reference-format YAML, then the first cycle must raise b200lp::Error —
    there is no CPU path behind the plugin interfaces."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    sc = synth.playground()
    yaml, scb = _write_case(str(tmp_path), sc.config, "differential_drive_simple", sc.cloud, sc.plan, sc.pose, sc.twist)
    p = _run(driver, yaml, scb, str(tmp_path / "out"), "differential_drive_simple", "early")
    assert p.returncode == 10, (p.returncode, p.stderr)
    assert "no usable CUDA device" in p.stderr and "no CPU fallback" in p.stderr


def test_unknown_plugin_type_is_an_error(driver, tmp_path):
    sc = synth.playground()
    cfg = PlannerConfig(generator=dict(sc.config.generator, plugin="trajectory_generators::NoSuchTheory"), critics=sc.config.critics)
    yaml, scb = _write_case(str(tmp_path), cfg, "g", sc.cloud, sc.plan, sc.pose, sc.twist)
    p = _run(driver, yaml, scb, str(tmp_path / "out"), "g", "early")
    assert p.returncode == 11 and "no plugin registered" in p.stderr


def _variant(generator, critics, twist, n_points=4000, seed=11):
    """small random scene around the origin with another generator / critic stack"""
    import copy
    import dataclasses
    sc = synth.playground()
    cfg = PlannerConfig(generator=copy.deepcopy(generator), critics=copy.deepcopy(critics))
    cloud = synth.to_xyzi(synth.small_scene(seed, n_points)[:, :3])
    return dataclasses.replace(sc, config=cfg, cloud=cloud, twist=list(twist))


CASES = {
    "playground": lambda: (synth.playground(), "differential_drive_simple", 0.0),
    "c1_ramp_small": lambda: (synth.c1_ramp(n_points=20_000), "differential_drive_simple", 0.0),
    "omni": lambda: (_variant(cfgmod.OMNI_SIMPLE_DEFAULT, cfgmod.OMNI_SIMPLE_CRITICS, (0.3, 0.1, 0.05)), "omni_drive_simple", 0.0),
    "rotate_shortest": lambda: (_variant(cfgmod.DD_ROTATE_INPLACE_DEFAULT, cfgmod.ROTATE_CRITICS, (0.0, 0.0, 0.0)),
                                "differential_drive_rotate_shortest_angle", -0.7),
}


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("mode", ["early", "late"])
def test_plugin_cycle_matches_oracle(driver, tmp_path, case, mode):
    from oracle import lporacle as O
    sc, gen, hdev = CASES[case]()
    yaml, scb = _write_case(str(tmp_path), sc.config, gen, sc.cloud, sc.plan, sc.pose, sc.twist, -1.0, hdev)
    prefix = str(tmp_path / "out")
    p = _run(driver, yaml, scb, prefix, gen, mode)
    assert p.returncode == 0, p.stderr
    s, traj, poses = _load(prefix)

    ora = O.OraclePlanner(sc.config, O.MATH_SHARED, O.INDEX_GRID)
    ora.set_cloud(sc.cloud)
    ora.set_plan(sc.plan)
    r = ora.plan(make_query(sc.pose, sc.twist, -1.0, hdev))
    t = ora.read_trajectories()

    # what the reference caller sees: list length, per-trajectory velocities / time_delta / summed cost, best pick
    assert int(s["n_traj"]) == r.n_traj == traj.shape[0]
    assert np.array_equal(traj[:, 0], t["cost"])                       # StackedScoringModel sum, bit for bit
    assert np.array_equal(traj[:, 1], t["vel"][:, 0].astype(np.float64))
    assert np.array_equal(traj[:, 3], t["vel"][:, 2].astype(np.float64))
    assert np.array_equal(traj[:, 4], t["time_delta"])
    assert np.array_equal(traj[:, 5].astype(np.int32), t["num_steps"])
    assert int(s["best_id"]) == r.best_id == int(s["device_best_id"]) == int(s["best_id2"])
    assert int(s["state"]) == int(s["state2"]) == (4 if r.best_id >= 0 else 2)  # TRAJECTORY_FOUND / ALL_TRAJECTORIES_FAIL
    if r.best_id >= 0:
        assert s["best_cost"] == r.best_cost and s["best_xv"] == r.xv and s["best_thetav"] == r.thetav
    # the patched caller needs one launch per cycle; the unpatched one re-launches when the critics bring the cloud
    # (a heading deviation first seen through the critics' shared data costs the first cycle one re-launch)
    assert int(s["launches_first"]) == (1 if mode == "early" and hdev == 0.0 else 2)
    assert int(s["launches_second"]) == (1 if mode == "early" else 2)

    # every pose / cuboid / AABB the Trajectory objects hold
    off = np.concatenate([[0], np.cumsum(t["num_steps"])])
    assert poses["pose"].shape[0] == off[-1]
    for tid in range(r.n_traj):
        po = ora.read_poses(tid, int(t["num_steps"][tid]))
        a, b = off[tid], off[tid + 1]
        for k in ("pose", "pcl_pose", "cuboid", "aabb"):
            assert np.array_equal(poses[k][a:b], po[k]), (case, tid, k)


@pytest.mark.gpu
@pytest.mark.parametrize("radius,expect_blocked", [(0.05, False), (2.5, True)])
def test_plugin_cycle_with_device_prune_plan_and_path_blocked_opinion(driver, tmp_path, radius, expect_blocked):
    """SURVEY.md §8(f) rows 1-2 through the host layer: Local_Planner::setPlan + prunePlan and the PathBlockedStrategy
    opinion, all on the device; the generator scores against the device-side prune plan (one launch per cycle)."""
    from oracle import lporacle as O
    sc = synth.c1_ramp(n_points=20_000)
    last, prev = sc.plan[-1], sc.plan[-2]
    ext = np.repeat(last[None, :], 80, 0)
    ext[:, :3] += (last[:3] - prev[:3])[None, :] * np.arange(1, 81)[:, None]
    gplan = np.concatenate([sc.plan, ext])
    gen = "differential_drive_simple"
    yaml, scb = _write_case(str(tmp_path), sc.config, gen, sc.cloud, gplan, sc.pose, sc.twist)
    prefix = str(tmp_path / "out")
    p = subprocess.run([driver, yaml, scb, prefix, gen, "early", "prune", "3.0", "1.0", str(radius)], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    s, traj, _ = _load(prefix)
    info, o_poses, o_pcl = O.prune_plan(gplan, sc.pose[:3], 3.0, 1.0)
    assert info.status == 0
    assert np.array_equal(np.fromfile(prefix + ".prune.f64", np.float64).reshape(-1, 7), o_poses)
    assert np.array_equal(np.fromfile(prefix + ".prunepcl.f32", np.float32).reshape(-1, 4), o_pcl)
    ora = O.OraclePlanner(sc.config, O.MATH_SHARED, O.INDEX_GRID)
    ora.set_cloud(sc.cloud)
    ora.set_plan(o_poses)
    r = ora.plan(make_query(sc.pose, sc.twist))
    assert np.array_equal(traj[:, 0], ora.read_trajectories()["cost"])
    assert int(s["best_id"]) == r.best_id
    b = ora.path_blocked(o_pcl, radius)
    assert s["blocked_ratio"] == b.ratio and int(s["blocked_opinion"]) == b.opinion == (1 if expect_blocked else 0)
    # PATH_BLOCKED_WAIT = 5 overrides the trajectory verdict (local_planner.cpp:597-607)
    assert int(s["state"]) == (5 if expect_blocked else (4 if r.best_id >= 0 else 2))
    assert int(s["launches_first"]) == 1 and int(s["launches_second"]) == 1


@pytest.mark.gpu
def test_plugin_cycle_fed_by_the_device_side_observation_producer(driver, tmp_path):
    """SURVEY.md §8(f) row 4 through the host layer: two MultiLayerSpinningLidar mirrors run cbSensor's filter chain on
    the device, StackedPerception::aggregateObservations concatenates there, and the cycle scores against that cloud with
    one launch and no cloud upload; the host copy in SharedData::aggregate_observation_ equals the oracle's restatement."""
    from oracle import lporacle as O
    sc = synth.playground()
    gen = "differential_drive_simple"
    yaml, scb = _write_case(str(tmp_path), sc.config, gen, np.zeros((0, 8), np.float32), sc.plan, sc.pose, sc.twist)
    scans, expect = [], []
    for seed, mount in ((21, 0.25), (22, -0.2)):
        scan, b2s, _ = synth.lidar_scan(n_beams=32, n_azimuth=1024, seed=seed, room=(16.0, 12.0, 2.5), n_pillars=12)
        b2s = b2s.copy()
        b2s[0] = mount
        b2s[2] -= 0.15
        scans.append((scan, b2s))
        expect.append(O.sensor_observation(scan, b2s, np.asarray(sc.pose, np.float64), 6.0, 1.2)[1])
    sb = str(tmp_path / "scans.bin")
    with open(sb, "wb") as f:
        f.write(struct.pack("<q", len(scans)))
        for scan, b2s in scans:
            f.write(struct.pack("<q", scan.shape[0]))
            f.write(np.asarray(b2s, np.float64).tobytes())
            f.write(np.ascontiguousarray(scan, np.float32).tobytes())
    prefix = str(tmp_path / "out")
    p = subprocess.run([driver, yaml, scb, prefix, gen, "early", "scan", sb, "6.0", "1.2"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    s, traj, _ = _load(prefix)
    host = np.concatenate(expect)
    obs = np.fromfile(prefix + ".obs.f32", np.float32).reshape(-1, 8)
    assert int(s["n_observation"]) == host.shape[0] and int(s["sensor0_n_points"]) == expect[0].shape[0]
    assert np.array_equal(obs[:, :4].view(np.uint32), host.view(np.uint32)) and not obs[:, 4:].any()
    ora = O.OraclePlanner(sc.config, O.MATH_SHARED, O.INDEX_GRID)
    ora.set_cloud(host)
    ora.set_plan(sc.plan)
    r = ora.plan(make_query(sc.pose, sc.twist))
    assert np.array_equal(traj[:, 0], ora.read_trajectories()["cost"])
    assert int(s["best_id"]) == r.best_id == int(s["best_id2"]) and 0 < r.n_collided < r.n_traj
    assert int(s["launches_first"]) == 1 and int(s["launches_second"]) == 1


def _yaw_of(x, y, z, w):
    """tf2::impl::getYaw, as recovery_behaviors::yaw_of restates it."""
    sqx, sqy, sqz, sqw = x * x, y * y, z * z, w * w
    sarg = -2 * (x * z - w * y) / (sqx + sqy + sqz + sqw)
    if sarg <= -0.99999:
        return -2 * math.atan2(y, x)
    if sarg >= 0.99999:
        return 2 * math.atan2(y, x)
    return math.atan2(2 * (x * y + w * z), sqw + sqx - sqy - sqz)


def _shortest_angular_distance(a, b):
    r = math.fmod((b - a) + math.pi, 2.0 * math.pi)
    return r + math.pi if r <= 0.0 else r - math.pi


@pytest.mark.gpu
@pytest.mark.parametrize("blocked", [False, True])
def test_rotate_inplace_behavior_loop_matches_the_restated_loop(driver, tmp_path, blocked):
    """SURVEY.md §8f row 3, the caller: RotateInPlaceBehavior's control loop (rotate_inplace_behavior.cpp:137-305) through
    the C++ mirror — the rotate-in-place theory and its critic stack on the device once per pass, a robot that turns by
    cmd.angular.z / frequency — against the same loop restated here around the oracle: per pass the selected trajectory,
    its cost (bit for bit), the published command, got_180, dist_left, the end of the loop and its RecoveryState; and the
    critics' cloud must be empty after every pass (:254-256). `blocked`: a wall next to the robot rejects every
    trajectory, the robot stands still and the behaviour gives up after 5 s (RECOVERY_FAIL)."""
    from oracle import lporacle as O
    sc, gen, hdev = CASES["rotate_shortest"]()
    cloud = sc.cloud
    if blocked:
        wall = synth.voxel_block(-0.3, 0.3, -0.3, 0.3, 0.2, 0.6)
        cloud = synth.to_xyzi(wall[:, :3])
    pose = [0.2, -0.1, 0.0, *synth.quat_from_rpy(0.0, 0.0, 0.4)]
    twist = [0.0, 0.0, 0.0]
    tol, freq, max_passes = 0.3, 10.0, 400
    yaml, scb = _write_case(str(tmp_path), sc.config, gen, cloud, sc.plan, pose, twist, -1.0, hdev)
    prefix = str(tmp_path / "out")
    p = subprocess.run([driver, yaml, scb, prefix, gen, "early", "rotate", str(max_passes), str(tol)], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    rows = np.fromfile(prefix + ".rotate.f64", np.float64).reshape(-1, 10)

    ora = O.OraclePlanner(sc.config, O.MATH_SHARED, O.INDEX_GRID)
    ora.set_cloud(cloud)
    ora.set_plan(sc.plan)
    yaw = start = current = _yaw_of(*pose[3:])
    got_180, now, last_valid = False, 0.0, 0.0
    expect = []
    for _ in range(max_passes):
        if not (not got_180 or abs(_shortest_angular_distance(current, start)) > tol):
            expect.append([yaw, 0.0, 0.0, 1.0, 3.0, -1.0, -1.0, float(got_180), 0.0, None])
            break
        q = [0.0, 0.0, math.sin(yaw / 2), math.cos(yaw / 2)]
        current = _yaw_of(*q)
        if not got_180:
            d180 = abs(_shortest_angular_distance(current, start + math.pi))
            dist_left = math.pi + d180
            got_180 = got_180 or d180 < tol
        else:
            dist_left = abs(_shortest_angular_distance(current, start))
        r = ora.plan(make_query(pose[:3] + q, twist, -1.0, hdev))
        row = [yaw, 0.0, 0.0, 0.0, 3.0, float(r.best_id), float(r.best_cost), float(got_180), dist_left, 0.0]
        if got_180:
            row[3] = 1.0
            expect.append(row)
            break
        if r.best_id < 0:
            if now - last_valid > 5.0:
                row[3], row[4] = 1.0, 4.0  # RECOVERY_FAIL
                expect.append(row)
                break
        else:
            row[1], row[2] = r.thetav, r.xv
            last_valid = now
        expect.append(row)
        yaw += row[1] / freq
        now += 1.0 / freq
    assert len(rows) == len(expect)
    for k, (got, want) in enumerate(zip(rows, expect)):
        for c in range(10):
            if want[c] is not None:
                assert got[c] == want[c], (k, c, got, want)
    last = rows[-1]
    if blocked:
        assert last[3] == 1.0 and last[4] == 4.0 and np.all(rows[:, 5] == -1) and 50 <= len(rows) <= 53
    else:
        assert last[3] == 1.0 and last[4] == 3.0 and last[7] == 1.0 and len(rows) > 50  # RECOVERY_DONE after the half turn
        assert np.all(rows[:-1, 1] != 0.0)
