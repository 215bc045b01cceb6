"""The host-side packing upload of b200lp_set_cloud (large PointXYZI / PointXYZ clouds are packed to 12 bytes per point by
host threads while earlier chunks cross PCIe) must be invisible in the results: same grid, same cycle as the same points
handed over as a tight (N,3) array, which is never packed."""
import numpy as np
import pytest

from dddmr_navigation_b200 import LocalPlanner, make_query, synth
from tests.helpers import assert_result_equal, assert_trajectories_equal


@pytest.mark.gpu
@pytest.mark.parametrize("n,cols", [(400_003, 8), (60_000, 8), (1_000_001, 4), (2_000_000, 8)])
def test_packed_upload_equals_tight_upload(n, cols):
    sc = synth.c2_dense(n_points=n, samples=(24.0, 24.0))
    xyz = np.ascontiguousarray(sc.cloud[:, :3])
    xyz[5] = np.nan  # non-finite points travel through the packing like any other
    xyz[n - 2, 1] = np.inf
    wide = np.zeros((n, cols), np.float32)
    wide[:, :3] = xyz
    wide[:, 3:] = 7.0  # padding / intensity bytes must not leak into the grid
    a, b = LocalPlanner(sc.config, device=0), LocalPlanner(sc.config, device=0)
    q = make_query(sc.pose, sc.twist)
    for rep in range(2):  # the second round reuses the thread pool and the staging buffer
        a.set_cloud(wide)
        b.set_cloud(xyz)
        ua, ub = a.last_upload(), b.last_upload()
        if n * cols * 4 >= (2 << 20):
            assert ua["pack_threads"] > 0 and ua["h2d_bytes"] == n * 12, ua
        else:
            assert ua["pack_threads"] == 0 and ua["h2d_bytes"] == n * cols * 4, ua
        assert ub == {"h2d_bytes": n * 12, "pack_threads": 0}
        assert a.grid_info() == b.grid_info()
        for p in (a, b):
            p.set_plan(sc.plan)
        ra, rb = a.plan(q), b.plan(q)
        assert_result_equal(ra, rb)
        assert_trajectories_equal(a.read_trajectories(), b.read_trajectories())
        assert ra.n_traj > 0
    a.close()
    b.close()


@pytest.mark.gpu
def test_pack_pool_survives_changing_cloud_sizes_and_small_clouds_bypass_it():
    sc = synth.c2_dense(n_points=600_000, samples=(16.0, 16.0))
    lp, ref = LocalPlanner(sc.config, device=0), LocalPlanner(sc.config, device=0)
    q = make_query(sc.pose, sc.twist)
    for m in (600_000, 1000, 70_001, 0, 600_000, 65_535):
        cloud = sc.cloud[:m]
        lp.set_cloud(cloud)
        ref.set_cloud(np.ascontiguousarray(cloud[:, :3]) if m else cloud)
        assert (lp.last_upload()["pack_threads"] > 0) == (m * 32 >= (2 << 20))
        assert lp.grid_info() == ref.grid_info()
        for p in (lp, ref):
            p.set_plan(sc.plan)
        assert_result_equal(lp.plan(q), ref.plan(q))
    lp.close()
    ref.close()


@pytest.mark.gpu
def test_pack_threads_environment_override(monkeypatch):
    """B200LP_PACK_THREADS is read when a ctx starts its pool: 0 keeps the raw copy, n pins the thread count."""
    sc = synth.c2_dense(n_points=300_000, samples=(12.0, 12.0))
    q = make_query(sc.pose, sc.twist)
    results = []
    for env, want_threads, want_bytes in (("0", 0, 300_000 * 32), ("3", 3, 300_000 * 12)):
        monkeypatch.setenv("B200LP_PACK_THREADS", env)
        lp = LocalPlanner(sc.config, device=0)
        lp.set_cloud(sc.cloud)
        assert lp.last_upload() == {"h2d_bytes": want_bytes, "pack_threads": want_threads}
        lp.set_plan(sc.plan)
        results.append((lp.plan(q), lp.read_trajectories(), lp.grid_info()))
        lp.close()
    assert_result_equal(results[0][0], results[1][0])
    assert_trajectories_equal(results[0][1], results[1][1])
    assert results[0][2] == results[1][2]
