"""The checking build (libb200lp_checks.so: the same sources with -DB200LP_CHECKS=1) runs the path's main shapes with a
device-side bounds assertion at every indexed store and list access of the grid build and the cycle. A failed assertion
traps, which the C ABI reports as a CUDA error — so these tests pass only if no index ever left its array. (The pool's
boxes do not allow compute-sanitizer; this build is the substitute.) Results are compared with the normal build."""
import os

import numpy as np
import pytest

from dddmr_navigation_b200 import LocalPlanner, abi, make_query, synth

pytestmark = pytest.mark.gpu


def _both(cfg):
    if not os.path.exists(abi.CHECKS_LIB_PATH):
        pytest.fail("libb200lp_checks.so is missing: run __graft_entry__.build()")
    return LocalPlanner(cfg), LocalPlanner(cfg, lib_path=abi.CHECKS_LIB_PATH)


def _same(a, b):
    assert a.as_dict() == b.as_dict()


@pytest.mark.parametrize("scene", ["playground", "c1", "c2_slice", "c3_slice"])
def test_checked_build_single_robot_cycles(scene):
    sc = {"playground": synth.playground, "c1": synth.c1_ramp,
          "c2_slice": lambda: synth.c2_dense(n_points=400_000), "c3_slice": lambda: synth.c3_multilevel(n_points=1_000_000)}[scene]()
    ref, chk = _both(sc.config)
    for lp in (ref, chk):
        lp.set_cloud(sc.cloud)
        lp.set_plan(sc.plan)
    for twist in (sc.twist, [0.1, 0.0, -0.4]):
        q = make_query(sc.pose, twist)
        _same(ref.plan(q), chk.plan(q))
        tr, tc = ref.read_trajectories(), chk.read_trajectories()
        for k in tr:
            assert np.array_equal(tr[k], tc[k], equal_nan=True), k
    ref.close()
    chk.close()


def test_checked_build_sample_shards_and_fleet():
    sc = synth.c1_ramp(n_points=50_000)
    ref, chk = _both(sc.config)
    for lp in (ref, chk):
        lp.set_cloud(sc.cloud)
        lp.set_plan(sc.plan)
    q = make_query(sc.pose, sc.twist)
    for count in (2, 5, 8):
        for rank in range(count):
            _same(ref.plan_shard(q, rank, count), chk.plan_shard(q, rank, count))
            assert ref.traj_count() == chk.traj_count()
    n = 24
    poses, twists, plans, offs = synth.fleet_queries(n, region=(-8.0, 8.0, -4.0, 4.0), levels=(0.0,), cloud=sc.cloud)
    qs = (abi.Query * n)()
    for i in range(n):
        qs[i] = make_query(poses[i], twists[i])
    plans, offs = np.ascontiguousarray(plans, np.float64), np.ascontiguousarray(offs, np.int64)
    for a, b in zip(ref.plan_batch(qs, plans, offs), chk.plan_batch(qs, plans, offs)):
        _same(a, b)
    ref.close()
    chk.close()


def test_a_failed_check_surfaces_as_an_error():
    """The checks are not decoration: with B200LP_CHECK_SELFTEST the checking build hands prep_kernel a sample count that is
    off by one, the kernel's LP_CHECK traps and the call fails with a CUDA error (in a process of its own: a trap poisons the
    CUDA context)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from dddmr_navigation_b200 import LocalPlanner, abi, make_query, synth\n"
        "sc = synth.playground()\n"
        "lp = LocalPlanner(sc.config, lib_path=abi.CHECKS_LIB_PATH)\n"
        "lp.set_cloud(sc.cloud); lp.set_plan(sc.plan)\n"
        "try:\n"
        "    lp.plan(make_query(sc.pose, sc.twist))\n"
        "except abi.B200LPError as e:\n"
        "    print('CAUGHT', e)\n"
        "    raise SystemExit(0)\n"
        "print('NO ERROR'); raise SystemExit(3)\n" % root)
    env = dict(os.environ, B200LP_CHECK_SELFTEST="1")
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0 and "CAUGHT" in p.stdout, (p.returncode, p.stdout[-500:], p.stderr[-500:])
    assert "b200lp check failed" in p.stdout + p.stderr
