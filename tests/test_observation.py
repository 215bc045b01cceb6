"""the observation producer in front of the path — MultiLayerSpinningLidar::cbSensor's
transform -> pass-through -> 0.1 m voxel filter -> transform (multilayer_spinning_lidar.cpp:232-269) and
StackedPerception::aggregateObservations (stacked_perception.cpp:128-140).

CPU part: the oracle restatement of the PCL filters against hand-derived cases (PCL itself is not vendored in the reference
tree, so this row's parity is UNPINNED beyond these). GPU part: the device pipeline (stable radix sort + ordered centroid
sums) against the oracle, bit for bit, and the observation feeding the plan cycle without leaving the device.

Parity bar of this row: voxel set, voxel order and per-voxel point counts identical; centroids bit-identical to the oracle
in its scan-order mode. Against PCL's own (unstable-sort) summation order a centroid may move by rounding only:
|delta| <= (k-1) * ulp(|sum|) / k per coordinate for a voxel of k points — measured in
test_summation_order_only_moves_centroids_by_rounding."""
import math
import os

import numpy as np
import pytest

from dddmr_navigation_b200 import LocalPlanner, abi, make_query, synth
from oracle import lporacle as O
from tests.helpers import assert_result_equal, assert_same_array, assert_trajectories_equal

ID7 = (0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0)
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rows(*pts):
    a = np.zeros((len(pts), 4), np.float32)
    a[:, :3] = np.asarray(pts, np.float32).reshape(-1, 3)
    a[:, 3] = 1.0
    return a


# --------------------------------------------------------------------------------------------------------------------
# CPU: the restatement against hand-derived cases
# --------------------------------------------------------------------------------------------------------------------
def test_points_of_one_voxel_become_their_float_mean_in_scan_order():
    scan = rows((0.51, 0.52, 0.53), (0.58, 0.56, 0.54), (0.55, 0.59, 0.51))
    info, obs = O.sensor_observation(scan, ID7, ID7, 5.0, 2.0)
    assert info.n_scan == 3 and info.n_window == 3 and info.n_points == 1
    s = np.zeros(3, np.float32)
    for p in scan[:, :3]:
        s = (s + p).astype(np.float32)
    assert_same_array(obs[0, :3], (s / np.float32(3.0)).astype(np.float32), "centroid")
    assert obs[0, 3] == 1.0


def test_pass_through_limits_are_inclusive_floats_and_drop_non_finite_points():
    w, h = 5.0, 1.5
    up = np.nextafter(np.float32(w), np.float32(np.inf))
    scan = rows((w, 0.05, 0.05), (up, 0.05, 0.05), (-w, 0.05, 0.05), (0.05, w, 0.05), (0.05, -up, 0.05), (0.05, 0.05, 0.0),
                (0.05, 0.05, -1e-7), (0.05, 0.05, h), (0.05, 0.05, np.nextafter(np.float32(h), np.float32(9))),
                (np.nan, 0.0, 0.5), (0.0, np.inf, 0.5), (0.0, 0.0, -np.inf))
    info, obs = O.sensor_observation(scan, ID7, ID7, w, h)
    assert info.n_window == 5 and info.n_points == 5  # x = +-w, y = w, z = 0 and z = h survive, each alone in its voxel
    kept = {tuple(float(v) for v in p) for p in obs[:, :3]}
    f = lambda *v: tuple(float(np.float32(c)) for c in v)
    assert kept == {f(w, 0.05, 0.05), f(-w, 0.05, 0.05), f(0.05, w, 0.05), f(0.05, 0.05, 0.0), f(0.05, 0.05, h)}


def test_voxels_leave_in_ascending_index_order_x_fastest_then_y_then_z():
    scan = rows((0.95, 0.05, 0.05), (0.05, 0.05, 0.95), (0.05, 0.95, 0.05), (0.05, 0.05, 0.05), (0.15, 0.05, 0.05))
    _, obs = O.sensor_observation(scan, ID7, ID7, 5.0, 2.0)
    order = [tuple(np.floor(p * 10).astype(int)) for p in obs[:, :3]]
    assert order == [(0, 0, 0), (1, 0, 0), (9, 0, 0), (0, 9, 0), (0, 0, 9)]


def test_transforms_are_double_affine_products_rounded_to_float_and_the_second_one_is_optional():
    yaw = math.radians(90.0)
    b2s = (1.0, 2.0, 0.5, 0.0, 0.0, math.sin(yaw / 2), math.cos(yaw / 2))
    g2b = (10.0, 20.0, 0.0, 0.0, 0.0, 0.0, 1.0)
    scan = rows((1.0, 0.0, 0.0))
    _, base = O.sensor_observation(scan, b2s, g2b, 5.0, 2.0, is_local_planner=False)
    _, glob = O.sensor_observation(scan, b2s, g2b, 5.0, 2.0, is_local_planner=True)
    # Rz(90) * (1,0,0) + t = (1, 3, 0.5) up to the rounding of Eigen's quaternion -> matrix expression
    assert np.allclose(base[0, :3], (1.0, 3.0, 0.5), atol=1e-6)
    assert np.allclose(glob[0, :3], (11.0, 23.0, 0.5), atol=1e-5)
    c, s = math.cos(yaw / 2), math.sin(yaw / 2)
    m00, m01 = 1.0 - (2 * s * s), -(2 * s) * c
    assert base[0, 0] == np.float32(m00 * 1.0 + m01 * 0.0 + 0.0 * 0.0 + 1.0)


def test_empty_scan_and_empty_window_give_an_empty_observation():
    info, obs = O.sensor_observation(np.zeros((0, 4), np.float32), ID7, ID7, 5.0, 2.0)
    assert info.n_points == 0 and len(obs) == 0
    info, obs = O.sensor_observation(rows((9.0, 0.0, 0.5), (0.0, 0.0, -0.5)), ID7, ID7, 5.0, 2.0)
    assert info.n_window == 0 and info.n_points == 0


def test_voxel_grid_leaves_the_cloud_alone_when_its_indices_would_overflow():
    """pcl::VoxelGrid::applyFilter: (dx*dy*dz) > INT32_MAX -> "Leaf size is too small for the input dataset", output = input.
    (The device path refuses such windows up front with B200LP_E_INVALID; see test_gpu_empty_scan_empty_window_and_error_paths.)"""
    scan = rows((-4000.0, -4000.0, 0.5), (4000.0, 4000.0, 90.0), (1.0, 1.0, 1.0), (1.0005, 1.0005, 1.0005))
    info, obs = O.sensor_observation(scan, ID7, ID7, 5000.0, 100.0, leaf=0.01, is_local_planner=False)
    assert info.n_window == 4 and info.n_points == 4
    assert_same_array(obs[:, :3], scan[:, :3], "unfiltered output")
    # the same four points with a leaf that fits: the two close ones merge
    info, obs = O.sensor_observation(scan, ID7, ID7, 5000.0, 100.0, leaf=10.0, is_local_planner=False)  # 800 x 800 x 9 voxels
    assert info.n_points == 3


def test_summation_order_only_moves_centroids_by_rounding():
    scan, b2s, g2b = synth.lidar_scan(n_beams=64, n_azimuth=2048)
    i0, a = O.sensor_observation(scan, b2s, g2b, 10.0, 2.0, is_local_planner=False, order_mode=0)
    i1, b = O.sensor_observation(scan, b2s, g2b, 10.0, 2.0, is_local_planner=False, order_mode=1)
    assert i0.as_dict() == i1.as_dict() and a.shape == b.shape
    # same voxel for every output row, coordinates within a few float ulps of a window coordinate (|x| <= 10 -> ulp 9.5e-7)
    assert np.abs(a - b).max() <= 4e-6
    assert 0 < (a != b).any(axis=1).sum() < len(a) // 2  # the order does matter for some voxels: hence the canonical order


def _golden_observation_cases():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg.OBSERVATION_CASES


@pytest.mark.parametrize("name", ["observation_16x512", "observation_24x700_base_frame"])
def test_oracle_reproduces_the_committed_observation_fixtures(name):
    """tests/golden/*.npz (made by tests/golden/make_golden.py) pin the restated filter chain against regressions."""
    kw, window, height, leaf, local = _golden_observation_cases()[name]
    want = np.load(os.path.join(GOLDEN, name + ".npz"))
    scan, b2s, g2b = synth.lidar_scan(**kw)
    info, obs = O.sensor_observation(scan, b2s, g2b, window, height, leaf, local)
    assert (info.n_scan, info.n_window, info.n_points) == (int(want["n_scan"]), int(want["n_window"]), int(want["n_points"]))
    assert_same_array(obs.view(np.uint32), want["observation"].view(np.uint32), "observation bits")


# --------------------------------------------------------------------------------------------------------------------
# GPU: the device pipeline against the oracle
# --------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def gpu():
    p = LocalPlanner(synth.playground().config, device=0)
    yield p
    p.close()


def check_against_oracle(gpu, scan, b2s, g2b, window, height, leaf=0.1, local=True, sensor=0):
    info = gpu.sensor_observation(sensor, scan, b2s, g2b, window, height, leaf, local)
    oinfo, oobs = O.sensor_observation(scan, b2s, g2b, window, height, leaf, local)
    for f in ("n_scan", "n_window", "n_points"):
        assert getattr(info, f) == getattr(oinfo, f), (f, getattr(info, f), getattr(oinfo, f))
    obs = gpu.read_observation(sensor, info.n_points)
    assert_same_array(obs.view(np.uint32), oobs.view(np.uint32), "observation bits")
    return info, obs


@pytest.mark.gpu
@pytest.mark.parametrize("n_beams,n_az", [(1, 1), (1, 5), (4, 1023), (4, 1024), (1, 4097), (16, 512), (32, 1024), (64, 2048), (128, 2048)])
def test_gpu_observation_is_bit_identical_to_the_oracle_over_scan_sizes(gpu, n_beams, n_az):
    scan, b2s, g2b = synth.lidar_scan(n_beams=max(n_beams, 1), n_azimuth=n_az, seed=synth.SEED0 + n_beams * 7 + n_az)
    info, _ = check_against_oracle(gpu, scan, b2s, g2b, 10.0, 2.0)
    assert info.n_launches > 0


@pytest.mark.gpu
def test_gpu_observation_of_a_stitched_scan_of_millions_of_points(gpu):
    """>= 2^21 points switch the sort passes to 4096-record tiles (16 records per thread)."""
    scan, b2s, g2b = synth.lidar_scan(n_beams=1100, n_azimuth=2048, seed=77)
    assert scan.shape[0] >= (1 << 21)
    check_against_oracle(gpu, scan, b2s, g2b, 10.0, 2.0)


@pytest.mark.gpu
@pytest.mark.parametrize("window,height,leaf,local", [(5.0, 1.5, 0.1, True), (10.0, 2.0, 0.1, False), (20.5, 2.5, 0.1, True),
                                                      (8.0, 2.0, 0.05, True), (8.0, 2.0, 0.25, False), (60.0, 3.0, 0.1, True),
                                                      (3.0, 0.4, 0.1, True)])
def test_gpu_observation_over_windows_leaves_and_pass_counts(gpu, window, height, leaf, local):
    scan, b2s, g2b = synth.lidar_scan(n_beams=48, n_azimuth=1500, room=(60.0, 40.0, 3.0), n_pillars=60)
    check_against_oracle(gpu, scan, b2s, g2b, window, height, leaf, local)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["observation_16x512", "observation_24x700_base_frame"])
def test_gpu_observation_reproduces_the_committed_fixtures(gpu, name):
    kw, window, height, leaf, local = _golden_observation_cases()[name]
    want = np.load(os.path.join(GOLDEN, name + ".npz"))
    scan, b2s, g2b = synth.lidar_scan(**kw)
    info = gpu.sensor_observation(5, scan, b2s, g2b, window, height, leaf, local)
    assert (info.n_scan, info.n_window, info.n_points) == (int(want["n_scan"]), int(want["n_window"]), int(want["n_points"]))
    assert_same_array(gpu.read_observation(5, info.n_points).view(np.uint32), want["observation"].view(np.uint32), "observation bits")


@pytest.mark.gpu
@pytest.mark.parametrize("cols", [3, 4, 8])
def test_gpu_observation_accepts_12_16_and_32_byte_strides(gpu, cols):
    scan, b2s, g2b = synth.lidar_scan(n_beams=16, n_azimuth=900)
    s = np.zeros((scan.shape[0], cols), np.float32)
    s[:, :3] = scan[:, :3]
    check_against_oracle(gpu, s, b2s, g2b, 10.0, 2.0)


@pytest.mark.gpu
def test_gpu_heavy_voxels_keep_scan_order(gpu):
    """60 000 points inside one 10 cm voxel and 20 000 spread out: the ordered float sum of the heavy voxel only comes out
    identical if the sort is stable over tiles, warps and passes."""
    rng = np.random.default_rng(5)
    heavy = rng.uniform(0.301, 0.399, (60_000, 3)).astype(np.float32) + np.float32([1.0, -2.0, 0.5])
    rest = rng.uniform(-6, 6, (20_000, 3)).astype(np.float32)
    rest[:, 2] = np.abs(rest[:, 2]) * 0.3
    pts = np.concatenate([heavy, rest])
    rng.shuffle(pts)
    scan = np.zeros((pts.shape[0], 4), np.float32)
    scan[:, :3] = pts
    info, obs = check_against_oracle(gpu, scan, ID7, ID7, 5.0, 2.0, local=False)
    assert info.n_window > 60_000


@pytest.mark.gpu
def test_gpu_empty_scan_empty_window_and_error_paths(gpu):
    info = gpu.sensor_observation(1, np.zeros((0, 4), np.float32), ID7, ID7, 5.0, 2.0)
    assert info.n_points == 0 and len(gpu.read_observation(1, 0)) == 0
    info = gpu.sensor_observation(1, rows((9.0, 0.0, 0.5), (0.0, 0.0, -0.5), (np.nan, 0, 0)), ID7, ID7, 5.0, 2.0)
    assert info.n_window == 0 and info.n_points == 0
    with pytest.raises(abi.B200LPError) as e:
        gpu.sensor_observation(0, rows((0, 0, 0)), ID7, ID7, 5000.0, 100.0)  # 1e5 x 1e5 x 1e3 voxels
    assert e.value.code == abi.E_INVALID
    with pytest.raises(abi.B200LPError) as e:
        gpu.sensor_observation(abi.MAX_SENSORS, rows((0, 0, 0)), ID7, ID7, 5.0, 2.0)
    assert e.value.code == abi.E_INVALID
    with pytest.raises(abi.B200LPError) as e:
        gpu.read_observation(7, 10)
    assert e.value.code == abi.E_STATE
    with pytest.raises(abi.B200LPError) as e:
        gpu.aggregate_observations([7])
    assert e.value.code == abi.E_STATE
    scan, b2s, g2b = synth.lidar_scan(n_beams=8, n_azimuth=256)
    info = gpu.sensor_observation(2, scan, b2s, g2b, 10.0, 2.0)
    with pytest.raises(abi.B200LPError) as e:
        gpu.read_observation(2, info.n_points - 1)
    assert e.value.code == abi.E_INVALID


@pytest.mark.gpu
def test_gpu_read_back_in_pointxyzi_layout(gpu):
    scan, b2s, g2b = synth.lidar_scan(n_beams=16, n_azimuth=700)
    info, obs16 = check_against_oracle(gpu, scan, b2s, g2b, 10.0, 2.0, sensor=3)
    obs32 = gpu.read_observation(3, info.n_points, stride=32)
    assert obs32.shape == (info.n_points, 8)
    assert_same_array(obs32[:, :4], obs16, "xyz1")
    assert not obs32[:, 4:].any()  # intensity 0 (pcl::copyPointCloud PointXYZ -> PointXYZI), padding 0


@pytest.mark.gpu
def test_gpu_aggregated_observations_feed_the_cycle_without_leaving_the_device():
    """Two sensors' observations concatenated on the device (aggregateObservations) == the same cloud uploaded from the
    host with set_cloud: same grid, same plan result, same trajectories, and the same as the oracle fed the host copy."""
    sc = synth.playground()
    a, b = LocalPlanner(sc.config, device=0), LocalPlanner(sc.config, device=0)
    ora = O.OraclePlanner(sc.config, O.MATH_SHARED, O.INDEX_GRID)
    # robot at the playground pose; scans taken in a room whose pillars land around the robot's plan
    pose = np.asarray(sc.pose, np.float64)
    g2b = pose.copy()
    clouds = []
    for s, (seed, mount) in enumerate(((21, 0.25), (22, -0.2))):
        scan, b2s, _ = synth.lidar_scan(n_beams=32, n_azimuth=1024, seed=seed, room=(16.0, 12.0, 2.5), n_pillars=12)
        b2s = b2s.copy()
        b2s[0] = mount
        b2s[2] -= 0.15  # the floor returns end up below base_link's z = 0 and are dropped by the z pass-through
        info = a.sensor_observation(s, scan, b2s, g2b, 6.0, 1.2)
        assert info.n_points > 100
        clouds.append(a.read_observation(s, info.n_points))
    total = a.aggregate_observations([0, 1])
    host = np.concatenate(clouds)
    assert total == host.shape[0]
    b.set_cloud(host)
    ora.set_cloud(host)
    assert a.grid_info() == b.grid_info()
    q = make_query(sc.pose, sc.twist)
    for p in (a, b, ora):
        p.set_plan(sc.plan)
    ra, rb, ro = a.plan(q), b.plan(q), ora.plan(q)
    assert_result_equal(ra, rb)
    assert_result_equal(ra, ro)
    assert_trajectories_equal(a.read_trajectories(), ora.read_trajectories())
    assert 0 < ra.n_collided < ra.n_traj and ra.best_id >= 0  # the observation is in the way of some trajectories, not all
    a.close()
    b.close()
