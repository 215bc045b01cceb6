"""Seeded synthetic workloads for the configurations BASELINE.json names.

Everything is generated with numpy from a fixed seed (20261018 + config number): lethal clouds as float32
``pcl::PointXYZI``-layout arrays (N,8) — x,y,z,pad,intensity,pad,pad,pad, 32-byte stride, exactly what
``SharedData::aggregate_observation_`` holds (perception_3d/include/perception_3d/shared_data.h:79) — robot
poses, twists and prune plans. The clouds hold obstacles only (no ground), like the reference's
``segmented_cloud_pure``; obstacle bases float 5 cm above the local ground so the z=0 cuboid face does not
trivially collide.
"""
from __future__ import annotations

import copy
import math
from dataclasses import dataclass, field

import numpy as np

from .config import DD_SIMPLE_CRITICS, DD_SIMPLE_DEFAULT, PlannerConfig

SEED0 = 20261018
VOX = 0.05


# ---------------------------------------------------------------------------------------------------
# small geometry helpers
# ---------------------------------------------------------------------------------------------------
def quat_from_rpy(roll: float, pitch: float, yaw: float):
    cr, sr = math.cos(roll / 2), math.sin(roll / 2)
    cp, sp = math.cos(pitch / 2), math.sin(pitch / 2)
    cy, sy = math.cos(yaw / 2), math.sin(yaw / 2)
    return (sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy,
            cr * cp * cy + sr * sp * sy)


def to_xyzi(xyz: np.ndarray) -> np.ndarray:
    """(N,3) -> (N,8) float32 in pcl::PointXYZI memory layout (32-byte stride)."""
    out = np.zeros((xyz.shape[0], 8), np.float32)
    out[:, :3] = xyz
    out[:, 3] = 1.0  # PCL's padding float of the xyz quadruple is 1.0
    return out


def voxel_block(x0, x1, y0, y1, z0, z1, filled=True, vox=VOX) -> np.ndarray:
    """5 cm voxel centres of an axis-aligned block; filled=False keeps a 2-voxel shell (walls, hollow boxes)."""
    xs = np.arange(x0 + vox / 2, x1, vox)
    ys = np.arange(y0 + vox / 2, y1, vox)
    zs = np.arange(z0 + vox / 2, z1, vox)
    if len(xs) == 0 or len(ys) == 0 or len(zs) == 0:
        return np.zeros((0, 3), np.float32)
    g = np.stack(np.meshgrid(xs, ys, zs, indexing="ij"), -1).reshape(-1, 3)
    if not filled:
        t = 2 * vox
        inner = ((g[:, 0] > x0 + t) & (g[:, 0] < x1 - t) & (g[:, 1] > y0 + t) & (g[:, 1] < y1 - t))
        g = g[~inner]
    return g.astype(np.float32)


@dataclass
class Scenario:
    name: str
    config: PlannerConfig
    cloud: np.ndarray            # (N,8) float32, PointXYZI layout
    pose: list                   # 7 doubles
    twist: list                  # vx, vy, wz
    plan: np.ndarray             # (M,7) float64
    extra_poses: list = field(default_factory=list)  # more (pose, twist, plan) triples on the same map


def _dd_config(**over) -> PlannerConfig:
    g = copy.deepcopy(DD_SIMPLE_DEFAULT)
    cub = over.pop("cuboid", None)
    g.update(over)
    if cub is not None:
        g["cuboid"] = cub
    return PlannerConfig(generator=g, critics=copy.deepcopy(DD_SIMPLE_CRITICS))


def big_cuboid():
    """1.2 x 0.8 x 1.0 m footprint of config C2: some corners are > 1 m from base_link, so the reference's
    d^2 < 1 radius pre-filter (collision_model.cpp:122) changes results and is exercised."""
    x0, x1, y0, y1, z0, z1 = -0.6, 0.6, -0.4, 0.4, 0.0, 1.0
    return {"flb": [x1, y1, z0], "frb": [x1, y0, z0], "flt": [x1, y1, z1], "frt": [x1, y0, z1],
            "blb": [x0, y1, z0], "brb": [x0, y0, z0], "blt": [x0, y1, z1], "brt": [x0, y0, z1]}


def _plan_polyline(start_xy, yaw0, n, step, curve_after, curve_deg, ground, pitch_of=None):
    """n plan poses, `step` apart, straight then curving by curve_deg in total; z from ground(x,y)."""
    pts = []
    x, y, yaw = start_xy[0], start_xy[1], yaw0
    dyaw = math.radians(curve_deg) / max(1, n - curve_after)
    for i in range(n):
        z = ground(x, y)
        pitch = pitch_of(x, y, yaw) if pitch_of else 0.0
        q = quat_from_rpy(0.0, pitch, yaw)
        pts.append([x, y, z, q[0], q[1], q[2], q[3]])
        if i >= curve_after:
            yaw += dyaw
        x += step * math.cos(yaw)
        y += step * math.sin(yaw)
    return np.asarray(pts, np.float64)


def _fill_boxes(rng, n_target, have, region, ground, keep_clear, size_xy=(0.3, 1.5), size_z=(0.3, 2.0),
                filled=True, max_boxes=100000):
    """Random boxes inside region=(x0,x1,y0,y1) until `n_target` points exist; keep_clear(x,y,r)->bool vetoes."""
    blocks, n = [], have
    x0, x1, y0, y1 = region
    for _ in range(max_boxes):
        if n >= n_target:
            break
        sx, sy = rng.uniform(*size_xy), rng.uniform(*size_xy)
        sz = rng.uniform(*size_z)
        cx, cy = rng.uniform(x0 + sx, x1 - sx), rng.uniform(y0 + sy, y1 - sy)
        if keep_clear(cx, cy, 0.5 * math.hypot(sx, sy)):
            continue
        zb = ground(cx, cy) + 0.05
        b = voxel_block(cx - sx / 2, cx + sx / 2, cy - sy / 2, cy + sy / 2, zb, zb + sz, filled=filled)
        blocks.append(b)
        n += len(b)
    return blocks


def _finish_cloud(rng, blocks, n_points):
    xyz = np.concatenate(blocks, 0)
    if len(xyz) > n_points:
        xyz = xyz[rng.permutation(len(xyz))[:n_points]]
    elif len(xyz) < n_points:  # top up with jittered copies (keeps the count exact)
        extra = xyz[rng.integers(0, len(xyz), n_points - len(xyz))] + rng.uniform(-0.02, 0.02, (n_points - len(xyz), 3)).astype(np.float32)
        xyz = np.concatenate([xyz, extra.astype(np.float32)], 0)
    xyz = xyz + rng.uniform(-0.01, 0.01, xyz.shape).astype(np.float32)  # de-grid: lidar points are not lattice points
    return to_xyzi(xyz.astype(np.float32))


# ---------------------------------------------------------------------------------------------------
# the playground fixture of the reference (the only deterministic scenario it ships)
# ---------------------------------------------------------------------------------------------------
def playground(goal=(3.0, 1.0)) -> Scenario:
    """local_planner_play_ground_node.cpp:206-233,280-287 with local_planner_play_ground.yaml:63-125."""
    cfg = _dd_config(sim_time=5.0)
    obst = np.array([[0.80, 0.60, 0.2], [0.75, 0.65, 0.2], [0.85, 0.55, 0.2], [0.70, 0.70, 0.2], [0.90, 0.50, 0.2]],
                    np.float32)
    plan = np.zeros((20, 7))
    plan[:, 6] = 1.0
    dx, dy = goal[0] / 20, goal[1] / 20
    for i in range(20):
        plan[i, 0], plan[i, 1] = dx * i, dy * i
    return Scenario("playground", cfg, to_xyzi(obst), [0, 0, 0, 0, 0, 0, 1], [0.4, 0.0, 0.0], plan)


# ---------------------------------------------------------------------------------------------------
# C1: DD simple + default critics on a 10 degree ramp, 200 k points, ~520 trajectories x <= 40 poses
# ---------------------------------------------------------------------------------------------------
def c1_ramp(n_points=200_000, seed=SEED0 + 1) -> Scenario:
    rng = np.random.default_rng(seed)
    slope = math.radians(10.0)
    # ramp rises along +x for x in [0, 25]; landings before and after
    def ground(x, y):
        return math.tan(slope) * min(max(x, 0.0), 25.0)
    cfg = _dd_config(linear_x_sample=20.0, angular_z_sample=25.0)
    rx, ry = 8.0, 0.0
    pose = [rx, ry, ground(rx, ry), *quat_from_rpy(0.0, -slope, 0.0)]  # nose up the ramp
    def clear(x, y, r):
        return math.hypot(x - rx, y - ry) < 1.3 + r
    blocks = [
        voxel_block(-7.0, 33.0, -10.0, -9.9, 0.05, 2.0 + math.tan(slope) * 25.0),   # side walls
        voxel_block(-7.0, 33.0, 9.9, 10.0, 0.05, 2.0 + math.tan(slope) * 25.0),
    ]
    # a few deterministic pillars inside the sampled fan so >= 30 % of the trajectories collide
    for (px, py, s) in [(9.9, 0.55, 0.3), (10.3, -0.75, 0.35), (9.6, 1.5, 0.4), (10.6, -0.1, 0.2), (9.4, -1.4, 0.3), (10.0, 0.0, 0.15)]:
        zb = ground(px, py) + 0.05
        blocks.append(voxel_block(px - s / 2, px + s / 2, py - s / 2, py + s / 2, zb, zb + 1.2))
    have = sum(len(b) for b in blocks)
    blocks += _fill_boxes(rng, n_points, have, (-7.0, 33.0, -9.8, 9.8), ground, clear, size_xy=(0.2, 0.9), size_z=(0.3, 1.5))
    cloud = _finish_cloud(rng, blocks, n_points)
    plan = _plan_polyline((rx - 0.5, ry), 0.0, 60, 0.05, 40, 20.0, ground,
                          pitch_of=lambda x, y, yaw: -slope if 0.0 <= x <= 25.0 else 0.0)
    return Scenario("C1", cfg, cloud, pose, [1.0, 0.0, 0.0], plan)


# ---------------------------------------------------------------------------------------------------
# C2: dense sampling, 3-D footprint, single 60 m x 60 m floor, 2 M points
# ---------------------------------------------------------------------------------------------------
def c2_dense(n_points=2_000_000, seed=SEED0 + 2, samples=(128.0, 128.0)) -> Scenario:
    rng = np.random.default_rng(seed)
    ground = lambda x, y: 0.0
    cfg = _dd_config(linear_x_sample=samples[0], angular_z_sample=samples[1], sim_time=3.0, cuboid=big_cuboid())
    pose = [0.0, 0.0, 0.0, *quat_from_rpy(0.0, 0.0, 0.0)]
    def clear(x, y, r):
        return math.hypot(x, y) < 1.6 + r
    H = 2.0
    blocks = [voxel_block(-30.0, 30.0, -30.0, -29.9, 0.05, H), voxel_block(-30.0, 30.0, 29.9, 30.0, 0.05, H),
              voxel_block(-30.0, -29.9, -30.0, 30.0, 0.05, H), voxel_block(29.9, 30.0, -30.0, 30.0, 0.05, H),
              # interior walls with a doorway the plan passes through
              voxel_block(3.4, 3.5, -30.0, -0.9, 0.05, H), voxel_block(3.4, 3.5, 1.1, 30.0, 0.05, H)]
    for (px, py, sx, sy, sz) in [(2.2, 1.3, 0.5, 0.5, 1.0), (2.6, -1.5, 0.6, 0.4, 0.8), (1.9, -0.95, 0.25, 0.25, 1.5),
                                 (2.9, 0.75, 0.3, 0.3, 0.4)]:
        blocks.append(voxel_block(px - sx / 2, px + sx / 2, py - sy / 2, py + sy / 2, 0.05, 0.05 + sz))
    have = sum(len(b) for b in blocks)
    blocks += _fill_boxes(rng, n_points, have, (-29.8, 29.8, -29.8, 29.8), ground, clear)
    cloud = _finish_cloud(rng, blocks, n_points)
    plan = _plan_polyline((-0.5, 0.0), 0.0, 80, 0.05, 50, 25.0, ground)
    return Scenario("C2", cfg, cloud, pose, [0.9, 0.0, 0.0], plan)


# ---------------------------------------------------------------------------------------------------
# C3: three 60 m x 45 m floors at z = 0/3/6 m joined by two 12 degree ramps, 8 M points
# ---------------------------------------------------------------------------------------------------
RAMP_DEG = 12.0


def _c3_ground():
    slope = math.tan(math.radians(RAMP_DEG))
    run = 3.0 / slope  # 14.1 m of ramp per floor
    # ramp A: floor 0 -> 1 along +x, x in [5, 5+run], y in [-3, 3]; ramp B: floor 1 -> 2 along -x, y in [8, 14]
    def ramp_a(x, y):
        return -3.0 <= y <= 3.0 and 5.0 <= x <= 5.0 + run
    def ramp_b(x, y):
        return 8.0 <= y <= 14.0 and 5.0 <= x <= 5.0 + run
    def ground(x, y, level=0):
        if level == 0 and ramp_a(x, y):
            return (x - 5.0) * slope
        if level == 1 and ramp_b(x, y):
            return 3.0 + (5.0 + run - x) * slope
        return 3.0 * level
    return ground, run, slope


def c3_multilevel(n_points=8_000_000, seed=SEED0 + 3, samples=(128.0, 128.0)) -> Scenario:
    rng = np.random.default_rng(seed)
    ground, run, slope = _c3_ground()
    ang = math.radians(RAMP_DEG)
    cfg = _dd_config(linear_x_sample=samples[0], angular_z_sample=samples[1], sim_time=3.0, cuboid=big_cuboid())
    X0, X1, Y0, Y1 = -30.0, 30.0, -22.5, 22.5
    # three robot stations: mid-floor 0, ramp A entry (floor 0 -> ramp), ramp A exit (ramp -> floor 1)
    stations = [
        ((-15.0, -10.0, 0.0), 0, 0.0),
        ((4.2, 0.0, 0.0), 0, 0.0),
        ((5.0 + run - 0.8, 0.0, 0.0), 0, -ang),
    ]
    robots = []
    for (sx, sy, yaw), level, pitch in stations:
        z = ground(sx, sy, level)
        robots.append(((sx, sy, z), level, pitch, yaw))
    def clear_all(x, y, r):
        for (sx, sy, _), _, _, _ in robots:
            if math.hypot(x - sx, y - sy) < 1.8 + r:
                return True
        # keep the ramps and their approaches drivable
        if -3.5 <= y <= 3.5 and 0.0 <= x <= 5.0 + run + 5.0:
            return True
        if 7.5 <= y <= 14.5 and 0.0 <= x <= 5.0 + run + 5.0:
            return True
        return False
    blocks = []
    per_floor = n_points // 3
    for level in range(3):
        zf = 3.0 * level
        fl = [voxel_block(X0, X1, Y0, Y0 + 0.1, zf + 0.05, zf + 2.0), voxel_block(X0, X1, Y1 - 0.1, Y1, zf + 0.05, zf + 2.0),
              voxel_block(X0, X0 + 0.1, Y0, Y1, zf + 0.05, zf + 2.0), voxel_block(X1 - 0.1, X1, Y0, Y1, zf + 0.05, zf + 2.0)]
        if level == 0:  # ramp A side rails + obstacles beside the ramp entry/exit
            for yy in (-3.2, 3.1):
                for xs in np.arange(5.0, 5.0 + run, 0.5):
                    zb = ground(float(xs) + 0.25, 0.0, 0) + 0.05
                    fl.append(voxel_block(float(xs), float(xs) + 0.5, yy, yy + 0.1, zb, zb + 1.0))
            for (px, py, s, h) in [(6.4, 1.35, 0.4, 1.2), (6.9, -1.5, 0.5, 0.9), (5.0 + run + 1.4, 1.2, 0.4, 1.0),
                                   (5.0 + run + 1.9, -1.45, 0.5, 1.3), (-13.0, -9.0, 0.5, 1.0), (-12.6, -11.4, 0.4, 1.4)]:
                zb = (3.0 if px > 5.0 + run else ground(px, 0.0, 0)) + 0.05
                fl.append(voxel_block(px - s / 2, px + s / 2, py - s / 2, py + s / 2, zb, zb + h))
        have = sum(len(b) for b in fl)
        g_level = (lambda lv: (lambda x, y: ground(x, y, lv)))(level)
        fl += _fill_boxes(rng, per_floor, have, (X0 + 0.2, X1 - 0.2, Y0 + 0.2, Y1 - 0.2), g_level, clear_all)
        blocks += fl
    cloud = _finish_cloud(rng, blocks, n_points)

    def plan_from(x, y, yaw, level, n=80):
        pts = []
        for i in range(n):
            lvl = level
            # crossing the top of ramp A puts the plan on floor 1
            z = ground(x, y, 0) if (level == 0 and x <= 5.0 + run) else (3.0 if level == 0 and x > 5.0 + run and -3 <= y <= 3 else ground(x, y, lvl))
            on_ramp = level == 0 and -3.0 <= y <= 3.0 and 5.0 <= x <= 5.0 + run
            q = quat_from_rpy(0.0, -ang if on_ramp else 0.0, yaw)
            pts.append([x, y, z, *q])
            x += 0.05 * math.cos(yaw)
            y += 0.05 * math.sin(yaw)
        return np.asarray(pts, np.float64)

    triples = []
    for (p, level, pitch, yaw) in robots:
        pose = [p[0], p[1], p[2], *quat_from_rpy(0.0, pitch, yaw)]
        triples.append((pose, [0.9, 0.0, 0.0], plan_from(p[0] - 0.5 * math.cos(yaw), p[1] - 0.5 * math.sin(yaw), yaw, level)))
    sc = Scenario("C3", cfg, cloud, triples[0][0], triples[0][1], triples[0][2], extra_poses=triples[1:])
    return sc


# ---------------------------------------------------------------------------------------------------
# C5: a fleet of robots on free ground of a map (poses, twists, 60-point plans), C1 sampling
# ---------------------------------------------------------------------------------------------------
def fleet_queries(n_robots: int, seed=SEED0 + 5, region=(-28.0, 28.0, -20.0, 20.0), levels=(0.0,), cloud=None, clearance=1.0):
    """-> (poses (n,7), twists (n,3), plans (n*60,7), offsets (n+1,)). Uniform over the region and the floors; with
    `cloud` given, robots are re-drawn until no lethal point lies within `clearance` metres (0.5 m occupancy cells per
    floor, a point counts for the floor whose ground it is at most 2.5 m above)."""
    rng = np.random.default_rng(seed)
    occ = None
    if cloud is not None:
        cs = 0.5
        nx, ny = int(math.ceil((region[1] - region[0]) / cs)) + 8, int(math.ceil((region[3] - region[2]) / cs)) + 8
        ox, oy = region[0] - 4 * cs, region[2] - 4 * cs
        occ = np.zeros((len(levels), nx, ny), bool)
        xyz = np.asarray(cloud)[:, :3]
        ix = np.floor((xyz[:, 0] - ox) / cs).astype(np.int64)
        iy = np.floor((xyz[:, 1] - oy) / cs).astype(np.int64)
        ok = (ix >= 0) & (ix < nx) & (iy >= 0) & (iy < ny)
        for li, zl in enumerate(levels):
            m = ok & (xyz[:, 2] >= zl) & (xyz[:, 2] < zl + 2.5)
            occ[li, ix[m], iy[m]] = True
        rad = int(math.ceil(clearance / cs))

        def free(x, y, li):
            cx, cy = int((x - ox) // cs), int((y - oy) // cs)
            return not occ[li, max(cx - rad, 0):cx + rad + 1, max(cy - rad, 0):cy + rad + 1].any()
    poses = np.zeros((n_robots, 7))
    twists = np.zeros((n_robots, 3))
    plans = np.zeros((n_robots * 60, 7))
    offs = np.arange(n_robots + 1, dtype=np.int64) * 60
    for i in range(n_robots):
        for _ in range(1000):
            x, y = rng.uniform(region[0], region[1]), rng.uniform(region[2], region[3])
            li = int(rng.integers(0, len(levels)))
            if occ is None or free(x, y, li):
                break
        z = float(levels[li])
        yaw = rng.uniform(-math.pi, math.pi)
        q = quat_from_rpy(0.0, 0.0, yaw)
        poses[i] = [x, y, z, *q]
        twists[i] = [rng.uniform(0.2, 1.0), 0.0, rng.uniform(-0.3, 0.3)]
        curve = math.radians(rng.uniform(-30.0, 30.0)) / 60
        px, py, pyaw = x - 0.3 * math.cos(yaw), y - 0.3 * math.sin(yaw), yaw
        for k in range(60):
            plans[i * 60 + k] = [px, py, z, *quat_from_rpy(0.0, 0.0, pyaw)]
            pyaw += curve
            px += 0.05 * math.cos(pyaw)
            py += 0.05 * math.sin(pyaw)
    return poses, twists, plans, offs


def small_scene(seed: int, n_points: int = 3000, extent: float = 4.0, cuboid=None, theory_cfg: PlannerConfig | None = None):
    """Random small cloud around the origin for fast parity tests (not a BASELINE config)."""
    rng = np.random.default_rng(seed)
    blocks = []
    have = 0
    while have < n_points:
        cx, cy = rng.uniform(-extent, extent, 2)
        if math.hypot(cx, cy) < 0.9:
            continue
        s = rng.uniform(0.1, 0.6)
        h = rng.uniform(0.2, 1.5)
        b = voxel_block(cx - s / 2, cx + s / 2, cy - s / 2, cy + s / 2, 0.05, 0.05 + h, filled=bool(rng.integers(0, 2)))
        blocks.append(b)
        have += len(b)
    return _finish_cloud(rng, blocks, n_points)


# ---------------------------------------------------------------------------------------------------
# lidar scans for the observation producer
# ---------------------------------------------------------------------------------------------------
def lidar_scan(n_beams: int = 32, n_azimuth: int = 1024, seed: int = SEED0 + 6, sensor_height: float = 0.6,
               room=(16.0, 12.0, 2.8), n_pillars: int = 12, dropout: float = 0.02, noise: float = 0.01):
    """One revolution of a multilayer spinning lidar in a box room with cylindrical pillars, as the (N,4) float32
    pcl::PointXYZ rows (sensor frame) that pcl::fromROSMsg hands to MultiLayerSpinningLidar::cbSensor
    (multilayer_spinning_lidar.cpp:178-183). Returns without a hit are NaN rows (a non-dense cloud), a few are +inf.
    -> (scan, base_from_sensor (7,), global_from_base (7,))."""
    rng = np.random.default_rng(seed)
    el = np.deg2rad(np.linspace(-25.0, 15.0, n_beams))
    az = np.linspace(-np.pi, np.pi, n_azimuth, endpoint=False)
    A, E = np.meshgrid(az, el)
    d = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], -1).reshape(-1, 3)
    hx, hy, hz = room[0] / 2, room[1] / 2, room[2]
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.full(d.shape[0], np.inf)
        for axis, lo, hi in ((0, -hx, hx), (1, -hy, hy), (2, -sensor_height, hz - sensor_height)):
            for plane in (lo, hi):
                tp = plane / d[:, axis]
                t = np.where((tp > 0) & (tp < t), tp, t)
        for _ in range(n_pillars):
            cx, cy = rng.uniform(-hx + 1, hx - 1), rng.uniform(-hy + 1, hy - 1)
            if math.hypot(cx, cy) < 1.0:
                continue
            r = rng.uniform(0.1, 0.4)
            a = d[:, 0] ** 2 + d[:, 1] ** 2
            b = -2 * (d[:, 0] * cx + d[:, 1] * cy)
            c = cx * cx + cy * cy - r * r
            disc = b * b - 4 * a * c
            tp = (-b - np.sqrt(np.where(disc >= 0, disc, np.nan))) / (2 * a)
            t = np.where((tp > 0) & (tp < t), tp, t)
    t = t + rng.normal(0.0, noise, t.shape)
    pts = (d * t[:, None]).astype(np.float32)
    drop = rng.random(pts.shape[0]) < dropout
    pts[drop] = np.nan
    pts[rng.random(pts.shape[0]) < dropout / 20] = np.inf
    scan = np.zeros((pts.shape[0], 4), np.float32)
    scan[:, :3] = pts
    scan[:, 3] = 1.0
    # sensor mounted 0.25 m ahead of base_link, sensor_height up, yawed 2 deg and rolled 1 deg; robot somewhere in the map
    b2s = np.array([0.25, 0.0, sensor_height, *quat_from_rpy(math.radians(1.0), 0.0, math.radians(2.0))], np.float64)
    g2b = np.array([12.5, -3.75, 0.02, *quat_from_rpy(0.0, math.radians(-3.0), math.radians(40.0))], np.float64)
    return scan, b2s, g2b
