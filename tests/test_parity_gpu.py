"""GPU parity: the sm_100a path (through the C ABI) against the CPU oracle on identical inputs.

Bar (BASELINE.json north_star): bit-exact num_steps, sample validity / trajectory-id mapping, collision
outcome (first colliding pose), selected trajectory id; per-critic float scores within 1e-4 relative.
Against the oracle's "shared" math mode the floats are required to be IDENTICAL (same functions, no FMA);
against its "libm" mode (glibc, what the reference binary calls) the 1e-4 tolerance applies.
"""
import copy
import math
import os

import numpy as np
import pytest

from dddmr_navigation_b200 import LocalPlanner, Local_Planner, PlannerConfig, PlannerState, abi, make_query, synth
from dddmr_navigation_b200.config import (DD_ROTATE_INPLACE_DEFAULT, OMNI_SIMPLE_CRITICS, OMNI_SIMPLE_DEFAULT,
                                          ROTATE_CRITICS)
from oracle import lporacle as O
from tests.helpers import (assert_result_equal, assert_same_array, assert_trajectories_equal, reference_argmin,
                           run_pair)

pytestmark = pytest.mark.gpu


def _pair(cfg, math_mode=O.MATH_SHARED, index_mode=O.INDEX_GRID):
    return LocalPlanner(cfg), O.OraclePlanner(cfg, math_mode, index_mode)


def _full_compare(gpu, ora, r_g, r_o, n_pose_trajs=6, exact=True):
    assert_result_equal(r_g, r_o, exact_cost=exact)
    tg, to = gpu.read_trajectories(), ora.read_trajectories()
    assert_trajectories_equal(tg, to, exact=exact)
    n = r_o.n_traj
    if n == 0:
        return tg
    ids = sorted(set(np.linspace(0, n - 1, min(n, n_pose_trajs)).astype(int).tolist()))
    for i in ids:
        ns = int(to["num_steps"][i])
        pg, po = gpu.read_poses(i, ns), ora.read_poses(i, ns)
        for k in po:
            if exact or k in ("collide", "n_r1"):
                assert_same_array(pg[k], po[k], f"traj {i} {k}")
            else:
                np.testing.assert_allclose(pg[k], po[k], rtol=1e-4, atol=1e-9, err_msg=f"traj {i} {k}")
    return tg


def test_playground_fixture_bit_exact():
    sc = synth.playground()
    gpu, ora = _pair(sc.config)
    r_g, r_o = run_pair(gpu, ora, sc.cloud, sc.plan, sc.pose, sc.twist)
    assert r_o.n_traj == 55 and r_o.n_poses == 2363
    _full_compare(gpu, ora, r_g, r_o, n_pose_trajs=55)


def test_playground_vs_libm_oracle():
    sc = synth.playground()
    gpu, ora = _pair(sc.config, O.MATH_LIBM, O.INDEX_BRUTE)
    r_g, r_o = run_pair(gpu, ora, sc.cloud, sc.plan, sc.pose, sc.twist)
    _full_compare(gpu, ora, r_g, r_o, exact=False)


def test_c1_ramp_full():
    sc = synth.c1_ramp()
    gpu, ora = _pair(sc.config)
    r_g, r_o = run_pair(gpu, ora, sc.cloud, sc.plan, sc.pose, sc.twist)
    assert r_o.n_traj == 520 and r_o.n_collided > 0.3 * 520
    tg = _full_compare(gpu, ora, r_g, r_o, n_pose_trajs=12)
    assert r_g.best_id == reference_argmin(tg["cost"])
    s_g, n_g = gpu.count_radius()
    s_o, n_o = ora.count_radius()
    assert (s_g, n_g) == (s_o, n_o)


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref (nanoflann build) not present")
def test_c1_ramp_vs_reference_nanoflann_libm():
    sc = synth.c1_ramp()
    gpu, ora = _pair(sc.config, O.MATH_LIBM, O.INDEX_NANOFLANN)
    r_g, r_o = run_pair(gpu, ora, sc.cloud, sc.plan, sc.pose, sc.twist)
    _full_compare(gpu, ora, r_g, r_o, n_pose_trajs=4, exact=False)


SMALL_CASES = []


def _cfg(gen_over=None, critics=None, gen_base=None):
    from dddmr_navigation_b200.config import DD_SIMPLE_CRITICS, DD_SIMPLE_DEFAULT
    g = copy.deepcopy(gen_base or DD_SIMPLE_DEFAULT)
    g.update(gen_over or {})
    return PlannerConfig(generator=g, critics=copy.deepcopy(critics if critics is not None else DD_SIMPLE_CRITICS))


def _straight_plan(n=30, step=0.1, yaw=0.0, z=0.0, start=(-0.3, 0.0)):
    q = synth.quat_from_rpy(0, 0, yaw)
    return np.array([[start[0] + i * step * math.cos(yaw), start[1] + i * step * math.sin(yaw), z, *q] for i in range(n)])


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_random_small_scenes_dd(seed):
    rng = np.random.default_rng(seed)
    cfg = _cfg({"linear_x_sample": 7.0, "angular_z_sample": 9.0, "sim_time": float(rng.uniform(1.0, 3.0))})
    cloud = synth.small_scene(seed)
    yaw = float(rng.uniform(-3.1, 3.1))
    pose = [float(rng.uniform(-0.3, 0.3)), float(rng.uniform(-0.3, 0.3)), 0.0, *synth.quat_from_rpy(0.02, -0.03, yaw)]
    twist = [float(rng.uniform(0.0, 1.0)), 0.0, float(rng.uniform(-0.5, 0.5))]
    gpu, ora = _pair(cfg)
    r_g, r_o = run_pair(gpu, ora, cloud, _straight_plan(yaw=yaw), pose, twist)
    _full_compare(gpu, ora, r_g, r_o)


def test_big_cuboid_radius_filter_matters():
    # corners of the 1.2 x 0.8 x 1.0 footprint are > 1 m from base_link: points inside the cuboid but with
    # d^2 >= 1 must NOT collide (collision_model.cpp:122)
    cfg = _cfg({"cuboid": synth.big_cuboid(), "linear_x_sample": 6.0, "angular_z_sample": 8.0, "sim_time": 2.0})
    cloud = synth.small_scene(11, n_points=6000, extent=3.0)
    gpu, ora = _pair(cfg)
    r_g, r_o = run_pair(gpu, ora, cloud, _straight_plan(), [0, 0, 0, 0, 0, 0, 1], [0.8, 0, 0.1])
    _full_compare(gpu, ora, r_g, r_o, n_pose_trajs=10)


def test_collision_min_max_and_reordered_stack():
    critics = [
        {"plugin": "mpc_critics::StickPathModel", "weight": 0.1},
        {"plugin": "mpc_critics::CollisionMinMaxModel", "weight": 1.0},
        {"plugin": "mpc_critics::TwirlingModel", "weight": 0.3},
        {"plugin": "mpc_critics::CollisionModel", "weight": 1.0},
        {"plugin": "mpc_critics::TowardGlobalPlanModel", "weight": 2.0},
        {"plugin": "mpc_critics::ShortestAngleModel", "weight": 0.5},
        {"plugin": "mpc_critics::PurePursuitModel", "translation_weight": 0.7, "orientation_weight": 0.2},
    ]
    cfg = _cfg({"linear_x_sample": 5.0, "angular_z_sample": 7.0}, critics)
    cloud = synth.small_scene(5, n_points=4000, extent=3.0)
    gpu, ora = _pair(cfg)
    r_g, r_o = run_pair(gpu, ora, cloud, _straight_plan(), [0, 0, 0, 0, 0, 0, 1], [0.5, 0, -0.1], heading_dev=-0.4)
    _full_compare(gpu, ora, r_g, r_o, n_pose_trajs=8)


def test_edge_cases_cloud_and_plan_sizes():
    cfg = _cfg({"linear_x_sample": 4.0, "angular_z_sample": 5.0})
    gpu, ora = _pair(cfg)
    four = synth.to_xyzi(np.array([[0.5, 0, 0.2]] * 4, np.float32))       # < 5 points: collision critic returns 0
    five = synth.to_xyzi(np.array([[0.5, 0, 0.2]] * 5, np.float32))
    nanpts = synth.to_xyzi(np.array([[np.nan, 0, 0.2]] * 3 + [[5, 5, 5]] * 3, np.float32))
    for cloud in (four, five, nanpts, synth.to_xyzi(np.zeros((0, 3), np.float32))):
        for plan in (_straight_plan(), _straight_plan(n=2), np.zeros((0, 7))):   # plan < 3 -> 10.0; empty -> -4
            r_g, r_o = run_pair(gpu, ora, cloud, plan, [0, 0, 0, 0, 0, 0, 1], [0.3, 0, 0.0])
            _full_compare(gpu, ora, r_g, r_o, n_pose_trajs=3)
    assert r_o.best_id == -1  # empty plan: pure pursuit rejects everything (pure_pursuit_model.cpp:62-64)


def test_speed_override_motor_constraint_and_degenerate_window():
    cfg = _cfg({"linear_x_sample": 6.0, "angular_z_sample": 6.0, "use_motor_constraint": True, "gear_ratio": 30.0,
                "max_motor_shaft_rpm": 2500.0})
    gpu, ora = _pair(cfg)
    cloud = synth.small_scene(7)
    for twist, ov in (([0.9, 0, 0.2], 0.3), ([0.02, 0, 0.0], -1.0), ([0.5, 0, -0.55], 0.45)):
        r_g, r_o = run_pair(gpu, ora, cloud, _straight_plan(), [0, 0, 0, 0, 0, 0, 1], twist, max_speed=ov)
        _full_compare(gpu, ora, r_g, r_o, n_pose_trajs=4)


@pytest.mark.parametrize("mode", ["0", "skew"])
def test_velocity_axes_kernel_chains_and_exception_lists_equal_the_host_closed_form(mode, monkeypatch):
    """Single-robot launches get their VelocityIterator axes as a closed form the host checked against the chains
    (PrepPlan, lp_kernels.cuh). B200LP_AXIS_DEBUG=0 makes the kernel run the chains itself; =skew hands it a closed
    form that is slightly off, so entries travel as exceptions or (more than 8 of them) force the kernel's own chains.
    Every path must give the oracle's samples bit for bit — dd with a zero crossing, omni with three axes."""
    monkeypatch.setenv("B200LP_AXIS_DEBUG", mode)
    sc = synth.c1_ramp(n_points=20_000)
    gpu, ora = _pair(sc.config)
    for twist in (sc.twist, [0.05, 0.0, -0.3], [0.6, 0.0, 0.0]):  # (a new window each time: the axis plan is rebuilt)
        r_g, r_o = run_pair(gpu, ora, sc.cloud, sc.plan, sc.pose, twist)
        _full_compare(gpu, ora, r_g, r_o, n_pose_trajs=3)
    gpu.close()
    cfg = PlannerConfig(generator=copy.deepcopy(OMNI_SIMPLE_DEFAULT), critics=copy.deepcopy(OMNI_SIMPLE_CRITICS))
    gpu, ora = _pair(cfg)
    cloud = synth.small_scene(9, n_points=4000)
    r_g, r_o = run_pair(gpu, ora, cloud, _straight_plan(), [0.1, -0.1, 0, *synth.quat_from_rpy(0, 0, 0.4)], [0.3, 0.1, 0.1])
    _full_compare(gpu, ora, r_g, r_o, n_pose_trajs=4)
    gpu.close()


def test_omni_theory():
    cfg = PlannerConfig(generator=copy.deepcopy(OMNI_SIMPLE_DEFAULT), critics=copy.deepcopy(OMNI_SIMPLE_CRITICS))
    gpu, ora = _pair(cfg)
    cloud = synth.small_scene(9, n_points=4000)
    r_g, r_o = run_pair(gpu, ora, cloud, _straight_plan(), [0.1, -0.1, 0, *synth.quat_from_rpy(0, 0, 0.4)], [0.3, 0.1, 0.1])
    assert r_o.n_traj > 100
    _full_compare(gpu, ora, r_g, r_o, n_pose_trajs=8)


def test_rotate_inplace_theory():
    cfg = PlannerConfig(generator=copy.deepcopy(DD_ROTATE_INPLACE_DEFAULT), critics=copy.deepcopy(ROTATE_CRITICS))
    gpu, ora = _pair(cfg)
    cloud = synth.small_scene(13, n_points=2000, extent=1.5)
    for hd in (0.5, -0.5):
        r_g, r_o = run_pair(gpu, ora, cloud, _straight_plan(), [0, 0, 0, 0, 0, 0, 1], [0.0, 0, 0.0], heading_dev=hd)
        assert r_o.n_traj == 2 and r_o.n_poses == 2 * 126
        _full_compare(gpu, ora, r_g, r_o, n_pose_trajs=2)


def test_tie_goes_to_last_index():
    # no cloud, no plan-dependent critics: symmetric +-w samples cost the same; `<=` keeps the LAST (local_planner.cpp:460)
    cfg = _cfg({"linear_x_sample": 3.0, "angular_z_sample": 4.0}, [{"plugin": "mpc_critics::TwirlingModel", "weight": 1.0}])
    gpu, ora = _pair(cfg)
    r_g, r_o = run_pair(gpu, ora, synth.to_xyzi(np.zeros((0, 3), np.float32)), _straight_plan(), [0, 0, 0, 0, 0, 0, 1], [0.5, 0, 0.0])
    t = ora.read_trajectories()
    zero_w = np.where(t["cost"] == t["cost"].min())[0]
    assert len(zero_w) > 1 and r_o.best_id == zero_w[-1]
    _full_compare(gpu, ora, r_g, r_o)


def test_reference_named_surface():
    sc = synth.playground()
    lp = Local_Planner(sc.config)
    lp.setObservation(sc.cloud)
    lp.setPlan(sc.plan)
    lp.setRobotState(sc.pose, sc.twist)
    state, best = lp.computeVelocityCommand("differential_drive_simple")
    assert state == PlannerState.TRAJECTORY_FOUND and best.id == 51 and best.xv_ == 0.5
    lp.setPlan(np.zeros((0, 7)))
    state, best = lp.computeVelocityCommand("differential_drive_simple")
    assert state == PlannerState.ALL_TRAJECTORIES_FAIL and best.cost_ == -1.0


def test_sample_shard_union_equals_whole():
    sc = synth.c1_ramp()
    gpu, ora = _pair(sc.config)
    r_g, r_o = run_pair(gpu, ora, sc.cloud, sc.plan, sc.pose, sc.twist)
    whole = gpu.read_trajectories()
    other = make_query(sc.pose, [0.35, 0.0, 0.25])  # a different window: whatever it leaves in the output arrays is wrong for `sc`
    for count in (2, 3, 8):
        best = (None, -1)
        cost = np.full(r_o.n_traj, np.nan)
        first_hit = np.full(r_o.n_traj, -7, np.int32)
        n_poses = n_traj = n_collided = 0
        ranges = []
        for rank in range(count):
            # the output arrays are not cleared between cycles: overwrite them with another query's values first, so that a
            # trajectory the shard launch skips cannot pass with what the unsharded run left there
            gpu.plan(other)
            r = gpu.plan_shard(make_query(sc.pose, sc.twist), rank, count)
            n_g, b, e = gpu.traj_count()
            assert n_g == r_o.n_traj
            ranges.append((b, e))
            t = gpu.read_trajectories()
            cost[b:e] = t["cost"][b:e]
            first_hit[b:e] = t["first_hit_pose"][b:e]
            assert np.all(np.isnan(t["cost"][:b])) and np.all(np.isnan(t["cost"][e:]))
            # the counters come from what plan_kernel scored, not from prep_kernel's list
            assert r.n_traj == e - b and r.n_poses == int(whole["num_steps"][b:e].sum())
            assert r.n_collided == int((whole["first_hit_pose"][b:e] >= 0).sum())
            n_poses += r.n_poses
            n_traj += r.n_traj
            n_collided += r.n_collided
            if r.best_id >= 0 and (best[0] is None or r.best_cost < best[0] or (r.best_cost == best[0] and r.best_id > best[1])):
                best = (r.best_cost, r.best_id)
        assert ranges[0][0] == 0 and ranges[-1][1] == r_o.n_traj and all(ranges[i][1] == ranges[i + 1][0] for i in range(count - 1))
        assert n_poses == r_o.n_poses and n_traj == r_o.n_traj and n_collided == r_o.n_collided
        assert_same_array(cost, whole["cost"], f"shard x{count} cost")
        assert_same_array(first_hit, whole["first_hit_pose"], f"shard x{count} first hit")
        assert best[1] == r_o.best_id and best[0] == r_o.best_cost
        # the cuts are balanced by estimated poses: no shard may hold a gross multiple of its share
        if count > 1:
            shares = [int(whole["num_steps"][b:e].sum()) for b, e in ranges]
            assert max(shares) <= 2.5 * r_o.n_poses / count + 64, shares


def test_fleet_batch_equals_individual_plans():
    sc = synth.c1_ramp(n_points=50_000)
    cfg = sc.config
    gpu = LocalPlanner(cfg)
    gpu.set_cloud(sc.cloud)
    n = 12
    poses, twists, plans, offs = synth.fleet_queries(n, region=(2.0, 20.0, -6.0, 6.0))
    qs = (abi.Query * n)()
    for i in range(n):
        qs[i] = make_query(poses[i], twists[i])
    res = gpu.plan_batch(qs, plans, offs)
    batch = [(res[i].as_dict(), gpu.read_trajectories(i)) for i in range(n)]
    ora = O.OraclePlanner(cfg)
    ora.set_cloud(sc.cloud)
    for i in range(n):
        ora.set_plan(plans[offs[i]:offs[i + 1]])
        r_o = ora.plan(qs[i])
        assert batch[i][0] == r_o.as_dict(), f"robot {i}"
        assert_trajectories_equal(batch[i][1], ora.read_trajectories())
    # a plan table in page-locked memory is uploaded from where it is (no staging copy): same results, and the staging
    # path still works afterwards
    import torch
    keep = torch.from_numpy(np.ascontiguousarray(plans, np.float64)).pin_memory()
    for table in (keep.numpy(), plans):
        res2 = gpu.plan_batch(qs, table, offs)
        assert [res2[i].as_dict() for i in range(n)] == [b[0] for b in batch]


def test_c2_full_size_properties_and_sampled_oracle():
    """BASELINE config C2 at full size (16.5 k trajectories, 2 M points): size-independent properties on the
    whole result + exact comparison with the oracle on every 64th sample."""
    sc = synth.c2_dense()
    gpu = LocalPlanner(sc.config)
    gpu.set_cloud(sc.cloud)
    gpu.set_plan(sc.plan)
    q = make_query(sc.pose, sc.twist)
    r = gpu.plan(q)
    t = gpu.read_trajectories()
    assert r.n_samples == 128 * 129 and r.n_traj == len(t["cost"])
    assert r.n_poses == int(t["num_steps"].sum())
    assert r.n_collided == int((t["first_hit_pose"] >= 0).sum())
    assert r.best_id == reference_argmin(t["cost"]) and r.best_cost == t["cost"][r.best_id]
    # idempotence: a second cycle on the same inputs returns the same bits
    r2 = gpu.plan(q)
    assert r2.as_dict() == r.as_dict()
    assert_same_array(gpu.read_trajectories()["cost"], t["cost"], "second cycle cost")
    stride = 64
    ora = O.OraclePlanner(sc.config)
    ora.set_cloud(sc.cloud)
    ora.set_plan(sc.plan)
    ora.set_sample_stride(stride)
    ora.plan(q, n_threads=8)
    to = ora.read_trajectories()
    sel = to["sample_index"]
    lut = {int(s): i for i, s in enumerate(t["sample_index"])}
    idx = np.array([lut[int(s)] for s in sel])
    for k in ("num_steps", "time_delta", "cost", "first_hit_pose", "critic_scores", "vel"):
        assert_same_array(t[k][idx], to[k], f"C2 sampled {k}")
    grid = gpu.grid_info()
    assert grid["n_points_kept"] == 2_000_000


def _sampled_oracle_check(sc, pose, twist, plan, gpu, stride, tag):
    q = make_query(pose, twist)
    gpu.set_plan(plan)
    r = gpu.plan(q)
    t = gpu.read_trajectories()
    assert r.n_poses == int(t["num_steps"].sum())
    assert r.n_collided == int((t["first_hit_pose"] >= 0).sum())
    assert r.best_id == reference_argmin(t["cost"])
    ora = O.OraclePlanner(sc.config)
    ora.set_cloud(sc.cloud)
    ora.set_plan(plan)
    ora.set_sample_stride(stride)
    ora.plan(q, n_threads=8)
    to = ora.read_trajectories()
    lut = {int(s): i for i, s in enumerate(t["sample_index"])}
    idx = np.array([lut[int(s)] for s in to["sample_index"]])
    for k in ("num_steps", "time_delta", "cost", "first_hit_pose", "critic_scores", "vel"):
        assert_same_array(t[k][idx], to[k], f"{tag} sampled {k}")
    return r, t


def test_c3_multilevel_full_size_all_stations():
    """BASELINE config C3 (8 M points, three floors joined by 12 degree ramps): mid-floor, ramp entry and ramp exit
    (the pitched start pose whose trajectories cross the floor transition), every 96th sample against the oracle."""
    sc = synth.c3_multilevel()
    gpu = LocalPlanner(sc.config)
    gpu.set_cloud(sc.cloud)
    assert gpu.grid_info()["n_points_kept"] == 8_000_000
    stations = [(sc.pose, sc.twist, sc.plan)] + list(sc.extra_poses)
    seen_hit = seen_free = False
    for i, (pose, twist, plan) in enumerate(stations):
        r, t = _sampled_oracle_check(sc, pose, twist, plan, gpu, 96, f"C3 station {i}")
        seen_hit |= r.n_collided > 0
        seen_free |= r.best_id >= 0
    assert seen_hit and seen_free


def test_c3_full_size_every_sample_against_the_oracle():
    """BASELINE config C3 at full size, NOT strided: the ramp-entry station (pitched start pose, trajectories crossing the
    floor transition) on the 8 M-point map, all 16.5 k trajectories — result, every per-trajectory array and every critic
    double — against the oracle on all host threads."""
    sc = synth.c3_multilevel()
    pose, twist, plan = sc.extra_poses[0]
    q = make_query(pose, twist)
    gpu = LocalPlanner(sc.config)
    gpu.set_cloud(sc.cloud)
    gpu.set_plan(plan)
    r = gpu.plan(q)
    ora = O.OraclePlanner(sc.config)
    ora.set_cloud(sc.cloud)
    ora.set_plan(plan)
    r_o = ora.plan(q, n_threads=os.cpu_count() or 8)
    assert r.n_traj > 16_000 and r.n_collided > 0
    assert r.as_dict() == r_o.as_dict()
    assert_trajectories_equal(gpu.read_trajectories(), ora.read_trajectories())


def test_c5_fleet_512_robots_against_single_cycles_and_the_oracle():
    """BASELINE config C5 at the per-GPU size the bench runs (512 robots, C1 sampling, the 8 M-point three-floor map), NOT
    sampled: EVERY robot of the batch against its own single-robot cycle and against the oracle (all of its samples, all host
    threads) — result struct field for field, every per-trajectory array bit for bit."""
    base = synth.c3_multilevel()
    cfg = synth.c1_ramp(n_points=1000).config
    n = 512
    poses, twists, plans, offs = synth.fleet_queries(n, region=(-28.0, 28.0, -20.0, 20.0), levels=(0.0, 3.0, 6.0), cloud=base.cloud)
    gpu = LocalPlanner(cfg)
    gpu.set_cloud(base.cloud)
    qs = (abi.Query * n)()
    for i in range(n):
        qs[i] = make_query(poses[i], twists[i])
    res = gpu.plan_batch(qs, plans, offs)
    batch = [(res[i].as_dict(), gpu.read_trajectories(i)) for i in range(n)]
    single = LocalPlanner(cfg)
    single.set_cloud(base.cloud)
    ora = O.OraclePlanner(cfg)
    ora.set_cloud(base.cloud)
    ora.set_keep_index(True)  # one map for all robots: the index is built once (the answers do not depend on it)
    threads = os.cpu_count() or 8
    found = 0
    for i in range(n):
        single.set_plan(plans[offs[i]:offs[i + 1]])
        r1 = single.plan(qs[i])
        assert r1.as_dict() == batch[i][0], f"robot {i}"
        assert_trajectories_equal(single.read_trajectories(), batch[i][1])
        found += r1.best_id >= 0
        ora.set_plan(plans[offs[i]:offs[i + 1]])
        r_o = ora.plan(qs[i], n_threads=threads)
        assert batch[i][0] == r_o.as_dict(), f"robot {i} vs oracle"
        assert_trajectories_equal(batch[i][1], ora.read_trajectories())
    assert found > n // 2


def test_c4_sample_shards_at_full_size():
    """BASELINE config C4 sampling (361 x 362 = 131 k samples) on the C3 map: the union of 8 contiguous sample shards
    reproduces the unsharded cycle bit for bit, and the reference argmin rule over the shard winners picks the same id."""
    sc = synth.c3_multilevel(n_points=2_000_000, samples=(361.0, 361.0))
    gpu = LocalPlanner(sc.config)
    gpu.set_cloud(sc.cloud)
    gpu.set_plan(sc.plan)
    q = make_query(sc.pose, sc.twist)
    r = gpu.plan(q)
    whole = gpu.read_trajectories()
    assert r.n_samples == 361 * 362  # the linear window stays positive (no inserted zero), the angular one straddles it
    from dddmr_navigation_b200.dist import cost_to_bits, pick_best
    count = 8
    cost = np.full(r.n_traj, np.nan)
    pairs, n_poses = [], 0
    for rank in range(count):
        rs = gpu.plan_shard(q, rank, count)
        _, b, e = gpu.traj_count()
        cost[b:e] = gpu.read_trajectories()["cost"][b:e]
        pairs.append((cost_to_bits(rs.best_cost, rs.best_id), rs.best_id))
        n_poses += rs.n_poses
    assert n_poses == r.n_poses
    assert_same_array(cost, whole["cost"], "C4 shard union cost")
    best_cost, best_id = pick_best(pairs)
    assert best_id == r.best_id and best_cost == r.best_cost


def test_c5_fleet_batch_on_the_multilevel_map():
    """BASELINE config C5 shape (C1 sampling, robots on all three floors of the shared map) at 96 robots: the batch equals
    per-robot cycles, and 8 of them are checked against the oracle in full."""
    base = synth.c3_multilevel(n_points=2_000_000)
    cfg = synth.c1_ramp(n_points=1000).config
    n = 96
    poses, twists, plans, offs = synth.fleet_queries(n, levels=(0.0, 3.0, 6.0), cloud=base.cloud)
    gpu = LocalPlanner(cfg)
    gpu.set_cloud(base.cloud)
    qs = (abi.Query * n)()
    for i in range(n):
        qs[i] = make_query(poses[i], twists[i])
    res = gpu.plan_batch(qs, plans, offs)
    batch = [(res[i].as_dict(), gpu.read_trajectories(i)) for i in range(n)]
    assert sum(1 for d, _ in batch if d["best_id"] >= 0) > n // 2  # robots stand on free ground
    single = LocalPlanner(cfg)
    single.set_cloud(base.cloud)
    ora = O.OraclePlanner(cfg)
    ora.set_cloud(base.cloud)
    for i in range(n):
        single.set_plan(plans[offs[i]:offs[i + 1]])
        r1 = single.plan(qs[i])
        assert r1.as_dict() == batch[i][0], f"robot {i}"
        assert_trajectories_equal(single.read_trajectories(), batch[i][1])
        if i % 12 == 0:
            ora.set_plan(plans[offs[i]:offs[i + 1]])
            r_o = ora.plan(qs[i])
            assert batch[i][0] == r_o.as_dict(), f"robot {i} vs oracle"
            assert_trajectories_equal(batch[i][1], ora.read_trajectories())


def _nccl_shard_worker(rank, world, port, out_q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from dddmr_navigation_b200.dist import allreduce_best
    sc = synth.c1_ramp()
    gpu = LocalPlanner(sc.config, device=rank)
    gpu.set_cloud(sc.cloud)
    gpu.set_plan(sc.plan)
    r = gpu.plan_shard(make_query(sc.pose, sc.twist), rank, world)
    cost, bid = allreduce_best(r.best_cost, r.best_id, device=torch.device("cuda", rank))
    out_q.put((rank, cost, bid, r.n_poses))
    dist.destroy_process_group()


def test_sample_sharding_over_nccl_two_gpus():
    """The real multi-GPU path of config C4: one process per GPU, plan_shard + the single NCCL all-reduce."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import multiprocessing as mp
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_shard_worker, args=(r, 2, port, out_q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(out_q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    sc = synth.c1_ramp()
    gpu, ora = _pair(sc.config)
    r_g, r_o = run_pair(gpu, ora, sc.cloud, sc.plan, sc.pose, sc.twist)
    assert got[0][1:3] == got[1][1:3] == (r_o.best_cost, r_o.best_id)
    assert got[0][3] + got[1][3] == r_o.n_poses


def _peer_exchange_worker(rank, world, device, q_in, q_out, n_cycles):
    """One process of a sample-sharded group whose argmin travels through peer device memory (no torch.distributed)."""
    sc = synth.c1_ramp()
    gpu = LocalPlanner(sc.config, device=device)
    gpu.set_cloud(sc.cloud)
    gpu.set_plan(sc.plan)
    q_out.put((rank, "handle", gpu.peer_export()))
    handles = q_in.get(timeout=300)
    gpu.peer_attach(rank, handles)
    out = []
    for c in range(n_cycles):
        tw = [sc.twist[0] - 0.05 * c, 0.0, 0.04 * c - 0.1]
        r = gpu.plan_shard_exchange(make_query(sc.pose, tw))
        out.append((r.best_id, r.best_cost, r.xv, r.yv, r.thetav, r.n_poses, r.n_traj))
    q_out.put((rank, "results", out))
    gpu.close()


def _run_peer_group(world, devices, n_cycles=6):
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    q_ins = [ctx.Queue() for _ in range(world)]
    procs = [ctx.Process(target=_peer_exchange_worker, args=(r, world, devices[r], q_ins[r], q_out, n_cycles)) for r in range(world)]
    for p in procs:
        p.start()
    handles = {}
    while len(handles) < world:
        rank, kind, payload = q_out.get(timeout=300)
        assert kind == "handle"
        handles[rank] = payload
    for qi in q_ins:
        qi.put([handles[r] for r in range(world)])
    results = {}
    while len(results) < world:
        rank, kind, payload = q_out.get(timeout=300)
        assert kind == "results"
        results[rank] = payload
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return results


def _check_peer_group(results, world, n_cycles=6):
    sc = synth.c1_ramp()
    gpu, ora = _pair(sc.config)
    for c in range(n_cycles):
        tw = [sc.twist[0] - 0.05 * c, 0.0, 0.04 * c - 0.1]
        r_g, r_o = run_pair(gpu, ora, sc.cloud, sc.plan, sc.pose, tw)
        want = (r_o.best_id, r_o.best_cost, r_o.xv, r_o.yv, r_o.thetav)
        for rank in range(world):
            assert results[rank][c][:5] == want, (c, rank, results[rank][c], want)  # every rank holds the GLOBAL best
        assert sum(results[rank][c][5] for rank in range(world)) == r_o.n_poses
        assert sum(results[rank][c][6] for rank in range(world)) == r_o.n_traj


def test_peer_memory_exchange_single_rank_equals_plan():
    sc = synth.c1_ramp()
    gpu, ora = _pair(sc.config)
    gpu.set_cloud(sc.cloud)
    gpu.set_plan(sc.plan)
    with pytest.raises(Exception):
        gpu.plan_shard_exchange(make_query(sc.pose, sc.twist))  # not attached yet
    gpu.peer_attach(0, [gpu.peer_export()])
    for c in range(4):
        tw = [sc.twist[0] - 0.1 * c, 0.0, 0.05 * c]
        q = make_query(sc.pose, tw)
        r_x = gpu.plan_shard_exchange(q)
        r_p = gpu.plan(q)
        assert r_x.as_dict() == r_p.as_dict()


def test_peer_memory_exchange_two_processes_on_one_gpu():
    """Two ranks of a sample-sharded group share GPU 0 (their kernels are time-sliced, so the in-kernel wait really waits):
    CUDA IPC mapping, sequence numbers and double buffering over six cycles; every rank ends up with the global best the
    oracle picks for the whole sample set."""
    _check_peer_group(_run_peer_group(2, [0, 0]), 2)


def test_peer_memory_exchange_over_nvlink_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _check_peer_group(_run_peer_group(2, [0, 1]), 2)


def _shared_cloud_clouds():
    """The clouds of the shared-cloud cycles: different sizes (also below the packing threshold and empty), so that piece
    boundaries, row-buffer reuse and the acknowledgements between cycles are all exercised."""
    sc = synth.c1_ramp()
    return sc, [sc.cloud, sc.cloud[:30_000], sc.cloud[:0], sc.cloud[50_000:180_001], sc.cloud]


def _shared_cloud_worker(rank, world, device, q_in, q_out):
    """One process of a peer group whose map is uploaded ONCE (by rank 0) and pushed to the others over peer memory."""
    sc, clouds = _shared_cloud_clouds()
    gpu = LocalPlanner(sc.config, device=device)
    gpu.set_plan(sc.plan)
    gpu.peer_reserve_cloud(max(len(c) for c in clouds))
    q_out.put((rank, "handle", gpu.peer_export()))
    gpu.peer_attach(rank, q_in.get(timeout=300))
    out = []
    q = make_query(sc.pose, sc.twist)
    for c, cloud in enumerate(clouds):
        gpu.set_cloud_shared(0, cloud if rank == 0 else None)
        r = gpu.plan(q)  # every rank scores the whole sample set on ITS copy of the map
        t = gpu.read_trajectories()
        out.append((r.as_dict(), t["cost"].tobytes(), t["first_hit_pose"].tobytes(), gpu.grid_info()["n_points_kept"],
                    gpu.last_upload()["h2d_bytes"]))
        # sample-sharded cycle on the shared map as well: both exchanges live in the same peer block
        x = gpu.plan_shard_exchange(q)
        out[-1] += ((x.best_id, x.best_cost),)
    q_out.put((rank, "results", out))
    gpu.close()


def _run_shared_cloud_group(world, devices):
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    q_ins = [ctx.Queue() for _ in range(world)]
    procs = [ctx.Process(target=_shared_cloud_worker, args=(r, world, devices[r], q_ins[r], q_out)) for r in range(world)]
    for p in procs:
        p.start()
    handles, results = {}, {}
    while len(handles) < world:
        rank, kind, payload = q_out.get(timeout=300)
        assert kind == "handle"
        handles[rank] = payload
    for qi in q_ins:
        qi.put([handles[r] for r in range(world)])
    while len(results) < world:
        rank, kind, payload = q_out.get(timeout=600)
        assert kind == "results"
        results[rank] = payload
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return results


def _check_shared_cloud_group(results, world):
    sc, clouds = _shared_cloud_clouds()
    gpu, ora = _pair(sc.config)
    for c, cloud in enumerate(clouds):
        r_g, r_o = run_pair(gpu, ora, cloud, sc.plan, sc.pose, sc.twist)
        t = gpu.read_trajectories()
        for rank in range(world):
            got = results[rank][c]
            assert got[0] == r_o.as_dict(), (c, rank)
            assert got[1] == t["cost"].tobytes() and got[2] == t["first_hit_pose"].tobytes(), (c, rank)
            assert got[3] == len(cloud), (c, rank)
            assert got[4] == (12 * len(cloud) if rank == 0 else 0), (c, rank)  # only the root crosses the host link
            assert got[5] == (r_o.best_id, r_o.best_cost), (c, rank)


def test_shared_cloud_two_processes_on_one_gpu():
    """The map of a peer group is uploaded by ONE rank and pushed to the other through peer memory (here both ranks share
    GPU 0: the same CUDA IPC mapping, flags and acknowledgements as over NVLink); every rank then plans on its own copy and
    must get what a plain set_cloud of the same cloud gives, bit for bit."""
    _check_shared_cloud_group(_run_shared_cloud_group(2, [0, 0]), 2)


def test_shared_cloud_over_nvlink_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _check_shared_cloud_group(_run_shared_cloud_group(2, [0, 1]), 2)


def test_first_cycle_of_a_fresh_process_is_already_right():
    """Regression: the very first launch in a process (lazy module load, cold caches, wide CTA start skew) once exposed
    a look-back poll the compiler had optimised away; later launches hid it."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from dddmr_navigation_b200 import LocalPlanner, make_query, synth\n"
            "sc = synth.c1_ramp(n_points=20_000)\n"
            "lp = LocalPlanner(sc.config); lp.set_cloud(sc.cloud); lp.set_plan(sc.plan)\n"
            "r = lp.plan(make_query(sc.pose, sc.twist))\n"
            "print('RESULT', r.n_samples, r.n_traj, r.n_poses, r.best_id, repr(r.best_cost))\n" % root)
    sc = synth.c1_ramp(n_points=20_000)
    ora = O.OraclePlanner(sc.config)
    ora.set_cloud(sc.cloud)
    ora.set_plan(sc.plan)
    r_o = ora.plan(make_query(sc.pose, sc.twist))
    for _ in range(3):
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr
        line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][0].split()
        assert [int(line[1]), int(line[2]), int(line[3]), int(line[4])] == [r_o.n_samples, r_o.n_traj, r_o.n_poses, r_o.best_id]
        assert float(line[5]) == r_o.best_cost


def test_api_edges_device_cloud_errors_and_reuse():
    """C-ABI behaviour around the path: a cloud that already lives on the device, 16-byte PointXYZ stride, call-order and
    argument errors (negative status + message, no crash), and one ctx reused across clouds of very different size."""
    import ctypes as C
    import torch
    sc = synth.c1_ramp(n_points=20_000)
    gpu, ora = _pair(sc.config)
    r_g, r_o = run_pair(gpu, ora, sc.cloud, sc.plan, sc.pose, sc.twist)
    # the same cloud handed over as a device buffer, PointXYZ (16-byte) layout
    xyz4 = np.ascontiguousarray(sc.cloud[:, :4])
    dev = torch.from_numpy(xyz4).cuda()
    gpu.set_cloud_device(dev.data_ptr(), dev.shape[0], 16)
    assert gpu.plan(make_query(sc.pose, sc.twist)).as_dict() == r_o.as_dict()
    # errors: bad stride, too long a plan, read-back before any cycle, out-of-range trajectory id
    lib = gpu.lib
    assert lib.b200lp_set_cloud(gpu.h, xyz4.ctypes.data_as(C.c_void_p), 10, 10) == abi.E_INVALID
    assert b"stride" in lib.b200lp_last_error(gpu.h)
    long_plan = np.zeros((abi.MAX_PLAN + 1, 7))
    assert lib.b200lp_set_plan(gpu.h, long_plan.ctypes.data_as(C.POINTER(C.c_double)), long_plan.shape[0]) == abi.E_INVALID
    fresh = LocalPlanner(sc.config)
    v = abi.TrajView()
    assert fresh.lib.b200lp_read_trajectories(fresh.h, 0, C.byref(v)) == abi.E_STATE
    pv = abi.PoseView()
    assert lib.b200lp_read_poses(gpu.h, 0, 10**6, C.byref(pv)) == abi.E_INVALID
    assert lib.b200lp_path_blocked(gpu.h, 0.5, C.byref(abi.Blocked())) == abi.E_STATE  # no device-side prune plan yet
    # the ctx survives the errors and a sequence of very different clouds (buffers grow and are reused)
    for n in (7, 200_000, 0, 3000, 20_000):
        cloud = synth.c1_ramp(n_points=max(n, 1000)).cloud[:n] if n else synth.to_xyzi(np.zeros((0, 3), np.float32))
        gpu.set_cloud(cloud)
        ora.set_cloud(cloud)
        q = make_query(sc.pose, sc.twist)
        assert gpu.plan(q).as_dict() == ora.plan(q).as_dict(), n
