"""CPU checks of bench.py's contract pieces that need no GPU: the reference arm's JSON line (thinned so that it ends in
seconds), its behaviour under a multi-rank launch (rank 0 alone works), and the shared-memory start gate of the
exchange steps with two processes."""
import json
import multiprocessing as mp
import os
import subprocess
import sys
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _run_reference(extra_env, *args):
    env = dict(os.environ)
    env.update(extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C1", "--steps", "1",
                           "--warmup", "0", "--ref-stride", "8", *args], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_prints_one_contract_line():
    p = _run_reference({})
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "trajectory_poses_scored_per_sec" and d["unit"] == "poses/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_under_torchrun_only_rank_zero_works():
    p = _run_reference({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2")
    assert p.returncode == 0, p.stderr[-2000:]
    assert p.stdout.strip() == ""


def _gate_worker(path, rank, world, rounds, delays, out):
    import bench
    g = bench.StartGate(path, rank, world)
    stamps = []
    for k in range(rounds):
        time.sleep(delays[rank] * (k % 3))  # ranks arrive at different times ...
        g.wait()
        stamps.append(time.perf_counter())  # ... and leave together
    out.put((rank, stamps))


def test_start_gate_releases_two_processes_together(tmp_path):
    import bench
    path = "/dev/shm/b200lp_test_gate_%d" % os.getpid()
    world, rounds = 2, 9
    bench.StartGate.create(path, world)
    try:
        ctx = mp.get_context("spawn")
        out = ctx.Queue()
        procs = [ctx.Process(target=_gate_worker, args=(path, r, world, rounds, (0.02, 0.05), out)) for r in range(world)]
        for p in procs:
            p.start()
        got = dict(out.get(timeout=120) for _ in range(world))
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:
        os.unlink(path)
    # perf_counter is CLOCK_MONOTONIC: comparable across processes of one host. The arrivals differ by up to 100 ms;
    # the departures must not (a loose bound: the test box is shared and the pollers are Python loops)
    skew = [abs(a - b) for a, b in zip(got[0], got[1])]
    assert max(skew) < 0.02, skew
    assert sorted(skew)[len(skew) // 2] < 0.002, skew


def test_start_gate_times_out_when_a_rank_is_missing():
    import bench
    path = "/dev/shm/b200lp_test_gate_t%d" % os.getpid()
    bench.StartGate.create(path, 2)
    try:
        g = bench.StartGate(path, 0, 2)
        with pytest.raises(RuntimeError):
            g.wait(timeout_s=0.2)
    finally:
        os.unlink(path)
