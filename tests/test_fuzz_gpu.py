"""Randomised parity sweep: many small scenes with random theory, critic stack, cuboid, tilted start pose, voxel-grid
cell size and a polluted cloud (NaN / inf points, far outliers), every per-trajectory output compared bit for bit with
the CPU oracle and the per-pose collision flags / n_r1 compared for a handful of trajectories per scene."""
import copy
import dataclasses
import math

import numpy as np
import pytest

from dddmr_navigation_b200 import LocalPlanner, PlannerConfig, make_query, synth
from dddmr_navigation_b200.config import (DD_ROTATE_INPLACE_DEFAULT, DD_SIMPLE_DEFAULT, OMNI_SIMPLE_DEFAULT)
from oracle import lporacle as O
from tests.helpers import assert_result_equal, assert_same_array, assert_trajectories_equal

pytestmark = pytest.mark.gpu

CRITIC_POOL = [
    {"plugin": "mpc_critics::CollisionModel", "weight": 1.0},
    {"plugin": "mpc_critics::CollisionMinMaxModel", "weight": 1.0},
    {"plugin": "mpc_critics::StickPathModel", "weight": 0.1},
    {"plugin": "mpc_critics::PurePursuitModel", "translation_weight": 1.0, "orientation_weight": 0.01},
    {"plugin": "mpc_critics::PurePursuitModel", "translation_weight": 0.3, "orientation_weight": 0.7},
    {"plugin": "mpc_critics::TowardGlobalPlanModel", "weight": 1.0},
    {"plugin": "mpc_critics::ShortestAngleModel", "weight": 0.5},
    {"plugin": "mpc_critics::TwirlingModel", "weight": 0.25},
]


def _random_cuboid(rng):
    # an axis-aligned box in the body frame, not centred, vertex order blb,brb,blt,flb,brt,frt,flt,frb
    xb, xf = -rng.uniform(0.1, 0.7), rng.uniform(0.1, 0.8)
    yr, yl = -rng.uniform(0.1, 0.6), rng.uniform(0.1, 0.6)
    zb, zt = rng.uniform(0.0, 0.1), rng.uniform(0.3, 1.3)
    v = {"blb": [xb, yl, zb], "brb": [xb, yr, zb], "blt": [xb, yl, zt], "flb": [xf, yl, zb], "brt": [xb, yr, zt],
         "frt": [xf, yr, zt], "flt": [xf, yl, zt], "frb": [xf, yr, zb]}
    return {k: [float(c) for c in p] for k, p in v.items()}


def _scene(seed):
    rng = np.random.default_rng(1000 + seed)
    theory = ["dd", "dd", "omni", "rotate"][seed % 4]
    if theory == "dd":
        gen = copy.deepcopy(DD_SIMPLE_DEFAULT)
        gen.update(linear_x_sample=float(rng.integers(3, 9)), angular_z_sample=float(rng.integers(3, 11)),
                   sim_time=float(rng.uniform(1.0, 3.5)), use_motor_constraint=bool(rng.integers(0, 2)),
                   max_vel_x=float(rng.choice([0.5, 1.0, 1.6])), acc_lim_x=float(rng.choice([0.3, 2.0])))
    elif theory == "omni":
        gen = copy.deepcopy(OMNI_SIMPLE_DEFAULT)
        gen.update(linear_x_sample=float(rng.integers(3, 6)), linear_y_sample=float(rng.integers(2, 5)),
                   angular_z_sample=float(rng.integers(3, 7)), sim_time=float(rng.uniform(1.0, 2.5)))
    else:
        gen = copy.deepcopy(DD_ROTATE_INPLACE_DEFAULT)
    gen["cuboid"] = _random_cuboid(rng)
    k = int(rng.integers(1, 7))
    critics = [copy.deepcopy(CRITIC_POOL[i]) for i in rng.choice(len(CRITIC_POOL), size=k, replace=True)]
    if seed % 4 != 3 and len(critics) < 8:  # most stacks hold a collision critic somewhere
        critics.insert(int(rng.integers(0, len(critics) + 1)), copy.deepcopy(CRITIC_POOL[int(rng.integers(0, 2))]))
    cfg = PlannerConfig(generator=gen, critics=critics, cell_xy=float(rng.choice([0.0, 0.1, 0.25, 0.5, 1.0])),
                        cell_z=float(rng.choice([0.0, 0.1, 0.4, 2.0])))
    cloud = synth.small_scene(500 + seed, n_points=int(rng.integers(200, 6000)), extent=float(rng.uniform(1.2, 3.0)))
    # pollution: non-finite points, far outliers (stretch the grid bounds), a point right at the robot
    extra = np.zeros((9, cloud.shape[1]), np.float32)
    extra[0, :3] = [np.nan, 0, 0]
    extra[1, :3] = [0, np.inf, 0]
    extra[2, :3] = [0, 0, -np.inf]
    extra[3, :3] = [rng.uniform(50, 300), rng.uniform(-300, -50), rng.uniform(-5, 40)]
    extra[4, :3] = [-rng.uniform(50, 300), rng.uniform(50, 300), rng.uniform(-20, 5)]
    extra[5:, :3] = rng.uniform(-1.0, 1.0, (4, 3))
    if rng.integers(0, 3):
        cloud = np.concatenate([cloud, extra[: int(rng.integers(1, 10))]])
    yaw = float(rng.uniform(-math.pi, math.pi))
    off = float(rng.choice([0.3, 1.0, 1.8]))  # the larger offsets put the robot among (sometimes inside) the obstacles
    pose = [float(rng.uniform(-off, off)), float(rng.uniform(-off, off)), float(rng.uniform(-0.2, 0.3)),
            *synth.quat_from_rpy(float(rng.uniform(-0.15, 0.15)), float(rng.uniform(-0.25, 0.25)), yaw)]
    twist = [float(rng.uniform(0.0, 1.3)), float(rng.uniform(-0.2, 0.2)), float(rng.uniform(-0.6, 0.6))]
    n_plan = int(rng.choice([0, 2, 3, 17, 60, 300]))
    pyaw = yaw + float(rng.uniform(-0.5, 0.5))
    plan = np.array([[pose[0] - 0.3 * math.cos(pyaw) + 0.05 * i * math.cos(pyaw), pose[1] - 0.3 * math.sin(pyaw) + 0.05 * i * math.sin(pyaw),
                      pose[2] + 0.002 * i, *synth.quat_from_rpy(0.0, float(rng.uniform(-0.1, 0.1)), pyaw)] for i in range(n_plan)]).reshape(-1, 7)
    if seed % 5 == 0:
        # the same scene far from the map origin: float coordinates get coarse (6e-5 m at 1 km), the conservative pre-test
        # and the candidate-box margins have to scale with the map extent
        off = np.array([rng.uniform(500, 3000), -rng.uniform(500, 3000), rng.uniform(-50, 120)])
        cloud = cloud.copy()
        fin = np.isfinite(cloud[:, :3]).all(axis=1)
        cloud[fin, :3] = (cloud[fin, :3].astype(np.float64) + off).astype(np.float32)
        pose = [pose[0] + off[0], pose[1] + off[1], pose[2] + off[2], *pose[3:]]
        if len(plan):
            plan = plan.copy()
            plan[:, :3] += off
    return cfg, cloud, plan, pose, twist, float(rng.choice([-1.0, 0.4])), float(rng.uniform(-1.0, 1.0))


@pytest.mark.parametrize("seed", range(160))
def test_fuzz_scene(seed):
    cfg, cloud, plan, pose, twist, max_speed, hdev = _scene(seed)
    gpu = LocalPlanner(cfg)
    ora = O.OraclePlanner(cfg, O.MATH_SHARED, O.INDEX_GRID if seed % 2 else O.INDEX_BRUTE)
    q = make_query(pose, twist, max_speed, hdev)
    for p in (gpu, ora):
        p.set_cloud(cloud)
        p.set_plan(plan)
    r_g, r_o = gpu.plan(q), ora.plan(q)
    assert_result_equal(r_g, r_o)
    t_g, t_o = gpu.read_trajectories(), ora.read_trajectories()
    assert_trajectories_equal(t_g, t_o)
    n = r_o.n_traj
    for tid in sorted(set([0, n // 3, n // 2, n - 1]) - {-1}):
        if tid < 0 or tid >= n:
            continue
        steps = int(t_o["num_steps"][tid])
        pg, po = gpu.read_poses(tid, steps), ora.read_poses(tid, steps)
        for k in po:
            assert_same_array(pg[k], po[k], f"seed {seed} traj {tid} {k}")
