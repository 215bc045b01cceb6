#!/usr/bin/env python
"""bench.py — trajectory-poses scored per second on the BASELINE.json workload.

    python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--workload C2|C1|C3|C4|C5]

A "step" is one local-plan cycle of the hot path (sampling -> rollout -> obstacle query -> critics -> argmin)
for one robot on the named synthetic workload (default C2: 128 x 129 = 16.5 k trajectories x <= 60 poses,
1.2 x 0.8 x 1.0 m footprint, 2 M-point single-floor lethal cloud). Metric: poses scored per second, where
poses = sum of num_steps over the generated trajectories (exactly what the oracle counts).

  value            device time of the cycle's kernels (CUDA events on the library's launching stream), cloud,
                   grid, plan and query resident in HBM; L2 flushed between steps.
  e2e              the same metric through the C ABI with HOST buffers: every step uploads the PointXYZI cloud
                   from pinned host memory, rebuilds the voxel grid (the reference rebuilds its kd-tree every
                   cycle, model_shared_data.h:78-81), uploads plan + query and reads the result back.
  roofline         fused plan kernel: algorithmic bytes (16 B x n_r1(pose) + 64 B per pose) / its
                   CUDA-event duration, against the measured HBM copy bandwidth (MEASURED_PEAKS.json).
  cpu_baseline     the reference's OWN theory/critic sources (oracle/_ref/liblpref.so, built from /root/reference against
                   stand-ins for its third-party headers), single thread as upstream, one or two full cycles of the
                   same workload on this box's host cores (N=1, rank 0 only); the oracle port when that .so is absent.

N > 1 (torchrun, one rank per GPU): fleet sharding, weak scaling — every rank plans for its own robot on its
own replica of the map; no collective on the data path. value = poses of all ranks / max-over-ranks time.
--workload C5 is the batched-fleet configuration (512 robots per GPU on the 8 M-point map, weak scaling, no
collective); --workload C4 is the sample-sharded one (131 k trajectories split over the ranks, strong scaling,
one NCCL all-reduce of 16*W bytes per cycle, step time = host-observed kernels + exchange).
--impl reference times the reference's own sources (same .so) on a thinned velocity sampling of the workload (rank 0 only).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "trajectory_poses_scored_per_sec"
UNIT = "poses/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(workload):
    """dram bytes per launch of plan_kernel from the committed ncu capture, if one exists for this workload."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get(workload, {}).get("plan_kernel_dram_bytes_per_launch")
    return None


FLEET_ROBOTS_PER_GPU = 512  # C5: 4096 robots over 8 GPUs


def make_workload(name: str, rank: int, world: int = 1):
    """-> dict(sc, desc, mode, and per-mode inputs). Modes: "single" (one robot, one plan() per step),
    "shard" (C4: one robot, sample grid split over the ranks, NCCL argmin exchange), "fleet" (C5: plan_batch)."""
    from dddmr_navigation_b200 import synth
    w = {"mode": "single"}
    if name == "C1":
        sc = synth.c1_ramp()
        desc = "C1: DD simple + default critics, 520 trajectories x <=40 poses, 200k-point 10deg ramp map"
    elif name == "C3":
        sc = synth.c3_multilevel()
        desc = "C3: 16.5k trajectories x <=60 poses, 1.2x0.8x1.0 m footprint, 8M-point 3-floor map with ramps"
    elif name == "C4":
        sc = synth.c3_multilevel(samples=(361.0, 361.0))
        desc = "C4: 131k trajectories (361x362 samples) x <=60 poses on the 8M-point 3-floor map, sample grid split contiguously over the ranks"
        w["mode"] = "shard"
    elif name == "C5":
        base = synth.c3_multilevel(samples=(20.0, 25.0))
        c1 = synth.c1_ramp(n_points=1000)
        sc = synth.Scenario("C5", c1.config, base.cloud, base.pose, base.twist, base.plan)
        n_total = FLEET_ROBOTS_PER_GPU * world
        poses, twists, plans, offs = synth.fleet_queries(n_total, region=(-28.0, 28.0, -20.0, 20.0), levels=(0.0, 3.0, 6.0),
                                                         cloud=base.cloud)
        lo, hi = rank * FLEET_ROBOTS_PER_GPU, (rank + 1) * FLEET_ROBOTS_PER_GPU
        w.update(mode="fleet", fleet=(poses[lo:hi], twists[lo:hi], plans[offs[lo]:offs[hi]], offs[lo:hi + 1] - offs[lo]))
        desc = (f"C5: fleet planning, {FLEET_ROBOTS_PER_GPU} robots per GPU ({n_total} in total) x 520 trajectories x <=40 poses "
                "on the shared 8M-point 3-floor map (replicated per GPU)")
    else:
        sc = synth.c2_dense()
        desc = "C2: 16.5k trajectories x <=60 poses, 1.2x0.8x1.0 m footprint, 2M-point single-floor lethal cloud"
    pose, twist, plan = list(sc.pose), list(sc.twist), sc.plan
    # N > 1 in "single" mode: every rank serves its own robot, and all robots issue the SAME query (the named
    # configuration), so the work per GPU is exactly the N=1 work — a clean weak-scaling measurement. Fleets of
    # different robots are the C5 workload.
    w.update(sc=sc, pose=pose, twist=twist, plan=plan, desc=desc)
    return w


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region: an NVML polling thread (10 ms period); falls back to the
    recipe's `nvidia-smi --query-gpu=... -lms` subprocess when pynvml is unavailable."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = f"/tmp/b200lp_clocks_{os.getpid()}.csv"
        self.thread = None
        self.samples = []

    def _nvml_loop(self, pynvml, handle):
        R = pynvml
        names = [("hw_slowdown", getattr(R, "nvmlClocksEventReasonHwSlowdown", getattr(R, "nvmlClocksThrottleReasonHwSlowdown", 0x8))),
                 ("hw_thermal_slowdown", getattr(R, "nvmlClocksEventReasonHwThermalSlowdown", getattr(R, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40))),
                 ("sw_thermal_slowdown", getattr(R, "nvmlClocksEventReasonSwThermalSlowdown", getattr(R, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20))),
                 ("sw_power_cap", getattr(R, "nvmlClocksEventReasonSwPowerCap", getattr(R, "nvmlClocksThrottleReasonSwPowerCap", 0x4)))]
        get_reasons = getattr(R, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(R, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop:
            try:
                sm = R.nvmlDeviceGetClockInfo(handle, R.NVML_CLOCK_SM)
                mask = int(get_reasons(handle))
                self.samples.append((float(sm), [n for n, bit in names if mask & int(bit)]))
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        try:
            import threading
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except Exception:
                    idx = self.gpu
            handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            self._stop = False
            self.thread = threading.Thread(target=self._nvml_loop, args=(pynvml, handle), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.thread is not None:
            self._stop = True
            self.thread.join(timeout=2)
            if self.samples:
                out.update(sm_mhz=statistics.median([s for s, _ in self.samples]), sm_max_mhz=self.sm_max, samples=len(self.samples),
                           reasons=sorted({r for _, rs in self.samples for r in rs}))
            return out
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx = [], []
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 8:
                    continue
                try:
                    sm.append(float(parts[0]))
                    mx.append(float(parts[1]))
                except ValueError:
                    continue
                for nm, val in zip(names, parts[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            os.remove(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def bind_to_gpu_numa_node(gpu_index: int):
    """Pin this process to the CPUs NVML calls ideal for its GPU, so that the pinned upload buffers are allocated on the
    GPU's own NUMA node (8 ranks uploading 64 MB each per step otherwise fight over one socket's memory and links)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[gpu_index]) if vis else gpu_index
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return sorted(os.sched_getaffinity(0))
    except Exception as e:  # affinity is an optimisation, never a requirement
        return f"unavailable ({type(e).__name__})"


def choose_pack_threads(numa, local_world: int) -> int:
    """Host threads each rank's library may use to pack the cloud for upload (B200LP_PACK_THREADS): 3/4 of this rank's share
    of the CPUs local to its GPU, at most 12, none below 4 (then the raw copy is as fast; with 8 ranks on a 32-CPU host the
    host's memory bandwidth is the limit and packing only adds traffic)."""
    cpus = len(numa) if isinstance(numa, list) and numa else (os.cpu_count() or 1)
    t = min(12, (cpus // max(1, local_world)) * 3 // 4)
    return t if t >= 4 else 0


def pinned_copy(arr: np.ndarray):
    """Page-locked host copy of arr (torch allocates; numpy views it)."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
    return t, t.numpy()


REF_SAMPLES = (64.0, 64.0)  # --impl reference: velocity samples per axis of the bounded sample (C2 itself is 128 x 128)


def run_reference(args, rank, world):
    """CPU arm. With oracle/_ref/liblpref.so present (built from the reference's OWN theory / critic sources, see
    oracle/Makefile) it times that code — single-threaded, as the reference is — on the named map and parameters with the
    velocity sampling thinned to REF_SAMPLES so that K steps end within minutes. Otherwise it times the oracle port on all
    host threads."""
    if rank != 0:
        return 0
    import dataclasses
    from dddmr_navigation_b200 import make_query
    from oracle import lporacle as O
    name = args.workload if args.workload in ("C1", "C2", "C3") else "C3"
    wl = make_workload(name, 0)
    sc, pose, twist, plan, desc = wl["sc"], wl["pose"], wl["twist"], wl["plan"], wl["desc"]
    q = make_query(pose, twist)
    if O.have_reference_sources():
        gen = dict(sc.config.generator)
        thinned = name != "C1"
        if thinned:
            gen.update(linear_x_sample=REF_SAMPLES[0], angular_z_sample=REF_SAMPLES[1])
        cfg = dataclasses.replace(sc.config, generator=gen)
        ref = O.ReferencePlanner(cfg)
        ref.set_plan(plan)
        threads, kind = 1, "reference"

        def step():
            ref.set_cloud(sc.cloud)  # the cloud object is new every cycle upstream; the kd-tree is rebuilt in updateData()
            return ref.plan(q)
        sample = ((f"the reference's own C++ (oracle/_ref/liblpref.so: theories, critics, stacked models compiled from /root/reference "
                   f"against stand-ins for Eigen/PCL/tf2/rclcpp, vendored nanoflann kd-tree), 1 thread as upstream; "
                   + (f"velocity sampling thinned from the workload's to {int(REF_SAMPLES[0])} x {int(REF_SAMPLES[1])} samples, "
                      if thinned else "the full workload, ")
                   + "kd-tree over the full cloud rebuilt every step"))
    else:
        use_ref = O.have_ref()
        ora = O.OraclePlanner(sc.config, O.MATH_LIBM, O.INDEX_NANOFLANN if use_ref else O.INDEX_GRID)
        ora.set_cloud(sc.cloud)
        ora.set_plan(plan)
        threads, kind = os.cpu_count() or 1, "port"
        ora.set_sample_stride(args.ref_stride)

        def step():
            return ora.plan(q, threads)
        sample = (f"oracle port, every {args.ref_stride}-th velocity sample, all host threads, kd-tree rebuilt every step; "
                  f"index={'reference-vendored nanoflann 1.5.1' if use_ref else 'oracle bucket grid'}, glibc libm")
    for _ in range(min(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    poses = 0
    for _ in range(args.steps):
        r = step()
        poses += r.n_poses
    dt = time.perf_counter() - t0
    value = poses / dt
    sample += f" ({r.n_traj} trajectories, {r.n_poses} poses per step)"
    base = {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
            "config": {"workload": desc, "threads": threads},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--ref-stride", type=int, default=1, help="--impl reference / cpu_baseline: score every k-th sample")
    ap.add_argument("--cpu-baseline-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--latency-cycles", type=int, default=1000)
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="C4 at N > 1: how the argmin crosses GPUs")
    ap.add_argument("--observation-scans", type=int, default=50, help="timed lidar scans through the observation producer (0: skip)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from dddmr_navigation_b200 import LocalPlanner, make_query

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank)  # before any pinned allocation: the upload buffers should sit next to the GPU
    if "B200LP_PACK_THREADS" not in os.environ:  # this process knows how many ranks share the host; the library does not
        os.environ["B200LP_PACK_THREADS"] = str(choose_pack_threads(numa, int(os.environ.get("LOCAL_WORLD_SIZE", world))))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its version banner to stdout when the communicator comes up (NCCL_DEBUG=VERSION / WARN); stdout belongs to
        # the one JSON line, so file descriptor 1 points at stderr until the first collective has run
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    wl = make_workload(args.workload, rank, world)
    sc, pose, twist, plan, desc, mode = wl["sc"], wl["pose"], wl["twist"], wl["plan"], wl["desc"], wl["mode"]
    pin_t, cloud = pinned_copy(sc.cloud)
    n_pts, stride = cloud.shape[0], cloud.shape[1] * 4
    plan = np.ascontiguousarray(plan, np.float64)
    q = make_query(pose, twist)
    lp = LocalPlanner(sc.config, device=local_rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def flush_l2():
        flush.zero_()
        torch.cuda.synchronize()

    if mode == "fleet":
        from dddmr_navigation_b200 import abi
        f_poses, f_twists, f_plans, f_offs = wl["fleet"]
        qs = (abi.Query * len(f_poses))()
        for i in range(len(f_poses)):
            qs[i] = make_query(f_poses[i], f_twists[i])
        f_plans = np.ascontiguousarray(f_plans, np.float64)
        f_offs = np.ascontiguousarray(f_offs, np.int64)

    # sample sharding over several GPUs: exchange the argmin through peer memory when the box allows CUDA IPC between the
    # ranks (--exchange nccl keeps the single NCCL all-reduce of 16*W bytes)
    peer_exchange = False
    if mode == "shard" and world > 1 and args.exchange == "peer":
        from dddmr_navigation_b200.dist import attach_peer_exchange
        peer_exchange = attach_peer_exchange(lp, device=torch.device("cuda", local_rank))

    def cycle():
        """One step of the hot path on resident inputs -> (poses scored on this rank, result summary)."""
        if mode == "fleet":
            res = lp.plan_batch(qs, f_plans, f_offs)
            return sum(int(r.n_poses) for r in res), res[0]
        if mode == "shard":
            if peer_exchange:  # the argmin travels through peer device memory inside the cycle's last kernel
                r = lp.plan_shard_exchange(q)
                return int(r.n_poses), r
            r = lp.plan_shard(q, rank, world)
            if world > 1:
                from dddmr_navigation_b200.dist import allreduce_best
                cost, bid = allreduce_best(r.best_cost, r.best_id, device=torch.device("cuda", local_rank))
                r.best_cost, r.best_id = cost, bid
            return int(r.n_poses), r
        r = lp.plan(q)
        return int(r.n_poses), r

    def upload():
        # returns as soon as the host buffer is consumed; the grid kernels run under the host work of the plan call
        lp.set_cloud_ptr(cloud.ctypes.data, n_pts, stride)
        if mode != "fleet":
            lp.set_plan(plan)

    # ---------------- device-resident arm: kernels only ----------------
    upload()
    grid_ms = lp.last_timing()["ms_grid_build"]  # (asking waits for the grid)
    for _ in range(args.warmup):
        flush_l2()
        poses_per_step, r = cycle()
    launches0 = lp.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    dev_ms, plan_k_ms, prep_k_ms, argmin_k_ms, wall_ms = [], [], [], [], []
    for _ in range(args.steps):
        flush_l2()
        t0 = time.perf_counter()
        poses_per_step, r = cycle()
        wall_ms.append(1e3 * (time.perf_counter() - t0))
        dev_ms.append(lp.last_timing()["ms_plan_kernels"])
        km = lp.last_kernel_ms()
        plan_k_ms.append(km["plan_kernel"])
        prep_k_ms.append(km["prep_kernel"])
        argmin_k_ms.append(km["argmin_kernel"])
    barrier()
    launches = lp.launch_count() - launches0
    # the sample-sharded cycle ends with a collective: its step time is the host-observed one (kernels + exchange)
    t_dev = (sum(wall_ms) if mode == "shard" and world > 1 else sum(dev_ms)) / 1e3

    # ---------------- e2e arm: host buffers through the C ABI, every step ----------------
    # warm-up of the host->device path: the first dozen pinned uploads of a process run at a fraction of the link rate
    # (measured: 0.37 ms vs 0.13 ms for the 6.4 MB C1 cloud, tools/ab_setcloud.py)
    for _ in range(max(12, args.warmup)):
        upload()
        cycle()
    barrier()
    e2e_ms, e2e_stage = [], {"ms_upload": 0.0, "ms_grid_build": 0.0, "ms_plan": 0.0}
    for _ in range(args.steps):
        flush_l2()
        t0 = time.perf_counter()
        upload()
        _, r2 = cycle()
        e2e_ms.append(1e3 * (time.perf_counter() - t0))
        tm = lp.last_timing()
        e2e_stage["ms_upload"] += tm["ms_upload"]
        e2e_stage["ms_grid_build"] += tm["ms_grid_build"]
        e2e_stage["ms_plan"] += tm["ms_plan_kernels"]
        assert (r2.best_id, r2.best_cost) == (r.best_id, r.best_cost)
    barrier()
    clocks = sampler.stop()
    t_e2e = sum(e2e_ms) / 1e3
    if mode == "fleet":
        h2d = lp.last_upload()["h2d_bytes"] + f_plans.nbytes + len(qs) * ctypes.sizeof(abi.Query)
        d2h = len(qs) * (56 + 32) + 32
    else:
        h2d = lp.last_upload()["h2d_bytes"] + plan.nbytes + ctypes.sizeof(q)
        d2h = 56 + 32 + 32  # result + meta + grid bounds

    # ---------------- reductions over ranks (max time, summed poses) ----------------
    poses_total = poses_per_step * args.steps
    if world > 1:
        t = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = float(t[0]), float(t[1])
        p = torch.tensor([poses_total], dtype=torch.int64, device="cuda")
        dist.all_reduce(p, op=dist.ReduceOp.SUM)
        poses_total = int(p[0])
    value = poses_total / t_dev
    e2e_value = poses_total / t_e2e

    # ---------------- roofline of the dominant kernel (rank 0's launch) ----------------
    sum_nr1, n_p = lp.count_radius()
    alg_bytes = 16 * sum_nr1 + 64 * n_p
    peak, peak_src = load_peaks()
    k_ms = sum(plan_k_ms) / len(plan_k_ms)
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": load_traffic(args.workload), "kernel": "plan_kernel", "kernel_ms": k_ms,
                "kernel_share_of_step": k_ms / (sum(dev_ms) / len(dev_ms)),
                "algorithmic_bytes_per_launch": alg_bytes, "sum_n_r1": sum_nr1, "poses": n_p, "peak_source": peak_src,
                "note": ("effective-bandwidth figure (SURVEY.md §8d): bytes the reference's radiusSearch(1.0) candidate "
                         "sets would stream; the voxel-grid prune touches far fewer and re-reads them from L1/L2, so the "
                         "kernel is issue/latency-bound, not DRAM-bound — see profiles/ for dram bytes and pipe utilisation")}

    gi = lp.grid_info()
    n_cells = gi["dims"][0] * gi["dims"][1] * gi["dims"][2]
    grid_bytes = n_pts * 48 + 4 * n_cells  # read the PointXYZI-stride input, write 16 B cell-sorted float4, offsets
    grid_ms_e2e = e2e_stage["ms_grid_build"] / args.steps
    roofline_grid = {"bound": "hbm", "achieved": grid_bytes / (grid_ms_e2e * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": grid_bytes / (grid_ms_e2e * 1e-3) / 1e9 / peak, "algorithmic_bytes": grid_bytes, "ms": grid_ms_e2e,
                     "note": ("whole grid build of one set_cloud (bounds, histogram, 3-kernel scan, scatter, two summed-volume passes, "
                              "two memsets and one host round trip for the grid dimensions); at this size it is launch/latency-bound, "
                              "and it sits behind a PCIe upload ~8x longer")}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "strong" if mode == "shard" else "weak",
        "vs_baseline": None,
        "dtype": "f32+f64", "data": "synthetic",
        "config": {"workload": desc, "trajectories": int(r.n_traj), "poses_per_step": poses_per_step, "cloud_points": n_pts,
                   "cloud_stride_bytes": stride, "timing": "CUDA events on the library stream; L2 flushed (256 MiB memset) between steps",
                   "parallelism": ({"single": "1 robot per GPU issuing the named query, map replicated (fleet sharding, no collective)",
                                    "fleet": f"{FLEET_ROBOTS_PER_GPU} robots per GPU, map replicated (fleet sharding, no collective)",
                                    "shard": ("sample grid split over the ranks, argmin exchanged through peer device memory over NVLink inside "
                                              "the cycle (exchange_kernel)" if peer_exchange else
                                              "sample grid split over the ranks, one 16*W-byte all-reduce per cycle (NCCL)")}[mode]
                                   if world > 1 or mode != "single" else "single GPU"),
                   "grid": lp.grid_info()},
        "cpu_affinity": (f"{len(numa)} CPUs local to the GPU: {numa[0]}-{numa[-1]}" if isinstance(numa, list) and numa else str(numa)),
        "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"]},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * t_e2e / args.steps,
                "stages_ms_per_step": {k: v / args.steps for k, v in e2e_stage.items()},
                "host_cloud_bytes_per_step": n_pts * stride, "host_pack_threads": lp.last_upload()["pack_threads"],
                "what": ("set_cloud (PointXYZI cloud in pinned host memory -> packed to 12 B/point by host_pack_threads host threads -> "
                         "upload + grid build) + set_plan + plan, host wall clock, every step")},
        "e2e_map_resident": {"value": poses_per_step * args.steps / (sum(wall_ms) / 1e3), "unit": UNIT,
                             "ms_per_step": sum(wall_ms) / len(wall_ms), "p50_ms": statistics.median(wall_ms),
                             "what": "b200lp_plan only (query+plan upload, kernels, result read-back), host wall clock, rank 0"},
        "gpu_launches": int(launches),
        "kernel_ms": {"prep_kernel": sum(prep_k_ms) / len(prep_k_ms), "plan_kernel": k_ms,
                      "argmin_kernel": sum(argmin_k_ms) / len(argmin_k_ms), "grid_build_total": grid_ms},
        "roofline": roofline,
        "roofline_grid_build": roofline_grid,
        "result": {"best_id": int(r.best_id), "best_cost": float(r.best_cost), "n_collided": int(r.n_collided)},
    }

    # ---------------- p50 cycle latency on the reference's own CPU-runnable case (C1), rank 0 ----------------
    if rank == 0 and world == 1 and args.latency_cycles > 0:
        try:  # an extra that fails must not cost the run its JSON line
            w1 = make_workload("C1", 0)
            sc1, pose1, twist1, plan1 = w1["sc"], w1["pose"], w1["twist"], w1["plan"]
            lp1 = LocalPlanner(sc1.config, device=local_rank)
            lp1.set_cloud(sc1.cloud)
            lp1.set_plan(plan1)
            rng = np.random.default_rng(0)
            lat = []
            for i in range(args.latency_cycles + 20):
                tw = [float(np.clip(twist1[0] + rng.uniform(-0.2, 0.0), 0.0, 1.0)), 0.0, float(rng.uniform(-0.2, 0.2))]
                q1 = make_query(pose1, tw)
                t0 = time.perf_counter()
                lp1.plan(q1)
                if i >= 20:
                    lat.append(1e3 * (time.perf_counter() - t0))
            line["p50_cycle_latency_ms"] = {"value": statistics.median(lat), "p99": float(np.percentile(lat, 99)),
                                            "cycles": len(lat), "workload": "C1 (520 trajectories, 200k-point map resident), perturbed twists"}
            lp1.close()
        except Exception as exc:  # noqa: BLE001
            line["p50_cycle_latency_ms"] = {"error": f"{type(exc).__name__}: {exc}"}

    # ---------------- the observation producer in front of the path, rank 0 ----------------
    if rank == 0 and world == 1 and args.observation_scans > 0:
        try:  # an extra that fails must not cost the run its JSON line
            from oracle import lporacle as O
            from dddmr_navigation_b200 import synth
            scan, b2s, g2b = synth.lidar_scan(n_beams=128, n_azimuth=2048)  # 262 144 points, one revolution
            _keep_scan, hscan = pinned_copy(scan)  # (torch tensor owning the pinned pages, numpy view)
            win, height = 10.0, 2.0
            for _ in range(10):
                oi = lp.sensor_observation(0, hscan, b2s, g2b, win, height)
            wall, dev, upl = [], [], []
            for _ in range(args.observation_scans):
                flush_l2()
                t0 = time.perf_counter()
                oi = lp.sensor_observation(0, hscan, b2s, g2b, win, height)
                wall.append(1e3 * (time.perf_counter() - t0))
                dev.append(oi.ms_device)
                upl.append(oi.ms_upload)
            passes = (oi.n_launches - 2) // 4
            n_s, n_w, n_o = int(oi.n_scan), int(oi.n_window), int(oi.n_points)
            obs_bytes = 32 * n_s + (16 * n_s + 16 * n_w) + (passes - 1) * 32 * n_w + 16 * n_w + 16 * n_o
            t0 = time.perf_counter()
            for _ in range(3):
                o_info, o_obs = O.sensor_observation(scan, b2s, g2b, win, height)
            cpu_ms = 1e3 * (time.perf_counter() - t0) / 3
            g_obs = lp.read_observation(0, n_o)
            d_ms = statistics.median(dev)
            line["observation"] = {
                "what": ("MultiLayerSpinningLidar::cbSensor filter chain (transform, 3 pass-throughs, 0.1 m voxel centroids, transform) on one "
                         "262 144-point scan from pinned host memory; the observation stays on the device"),
                "scan_points": n_s, "window_points": n_w, "observation_points": n_o, "radix_passes": passes, "launches": int(oi.n_launches),
                "ms_device_p50": d_ms, "ms_upload_p50": statistics.median(upl), "ms_host_wall_p50": statistics.median(wall), "scan_points_per_sec_e2e": n_s / (statistics.median(wall) * 1e-3),
                "roofline": {"bound": "hbm", "achieved": obs_bytes / (d_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": obs_bytes / (d_ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": obs_bytes,
                             "note": "upload + ~10 dependent launches on 4 MB of records: launch/latency-bound, not HBM-bound"},
                "cpu_baseline": {"ms": cpu_ms, "kind": "port", "cores": 1,
                                 "sample": "3 runs of the oracle restatement of the PCL filter chain on the same scan"},
                "matches_oracle_bits": bool(np.array_equal(g_obs.view(np.uint32), o_obs.view(np.uint32))),
            }
            del _keep_scan
        except Exception as exc:  # noqa: BLE001
            line["observation"] = {"error": f"{type(exc).__name__}: {exc}"}

    # ---------------- CPU baseline on this box's host cores (rank 0, N=1) ----------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline and mode == "single":
        try:  # an extra that fails must not cost the run its JSON line
            from oracle import lporacle as O
            if O.have_reference_sources() and max(1, args.ref_stride) == 1:
                # the reference's own sources, the FULL workload, one thread (as upstream); ~10 s per cycle at C2
                ref = O.ReferencePlanner(sc.config)
                ref.set_plan(plan)
                n_cycles = max(1, min(args.cpu_baseline_steps, 2))
                t0 = time.perf_counter()
                cp = 0
                for _ in range(n_cycles):
                    ref.set_cloud(sc.cloud)
                    ro = ref.plan(q)
                    cp += ro.n_poses
                dt = time.perf_counter() - t0
                assert ro.best_id == r.best_id, (ro.best_id, r.best_id)  # the GPU picks the trajectory the reference's own code picks
                line["cpu_baseline"] = {
                    "value": cp / dt, "unit": UNIT, "cores": 1, "kind": "reference",
                    "sample": (f"{n_cycles} full cycle(s) of the same workload ({ro.n_poses} poses each) through the reference's own C++ "
                               "(oracle/_ref/liblpref.so: its theory / critic / stacked-model sources compiled from /root/reference against "
                               "stand-ins for Eigen/PCL/tf2/rclcpp, its vendored nanoflann as kd-tree), kd-tree rebuilt every cycle, "
                               f"1 thread (the reference path is single-threaded); host has {os.cpu_count()} cores"),
                    "ms_per_cycle": 1e3 * dt / n_cycles,
                    "best_id_matches_gpu": bool(ro.best_id == r.best_id), "best_cost_matches_gpu_1e-4": bool(abs(ro.best_cost - r.best_cost) <= 1e-4 * abs(ro.best_cost)),
                }
                # courtesy upper bound: the oracle port with the trajectories split over every host core
                # (the reference itself is single-threaded; the kd-tree build stays serial)
                ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
                ora = O.OraclePlanner(sc.config, O.MATH_LIBM, O.INDEX_NANOFLANN if O.have_ref() else O.INDEX_GRID)
                ora.set_plan(plan)
                t0 = time.perf_counter()
                ora.set_cloud(sc.cloud)
                rm = ora.plan(q, ncores)
                dtm = time.perf_counter() - t0
                line["cpu_baseline"]["all_cores"] = {
                    "value": rm.n_poses / dtm, "unit": UNIT, "cores": ncores, "kind": "port", "ms_per_cycle": 1e3 * dtm,
                    "sample": "1 full cycle of the same workload through the oracle port, trajectories split over std::threads, index rebuilt (serial)",
                    "best_id_matches_gpu": bool(rm.best_id == r.best_id)}
            else:
                use_ref = O.have_ref()
                ora = O.OraclePlanner(sc.config, O.MATH_LIBM, O.INDEX_NANOFLANN if use_ref else O.INDEX_GRID)
                ora.set_cloud(sc.cloud)
                ora.set_plan(plan)
                stride_s = max(1, args.ref_stride)
                ora.set_sample_stride(stride_s)
                t0 = time.perf_counter()
                cp = 0
                for _ in range(args.cpu_baseline_steps):
                    ro = ora.plan(q, 1)
                    cp += ro.n_poses
                dt = time.perf_counter() - t0
                if stride_s == 1:
                    assert ro.best_id == r.best_id, (ro.best_id, r.best_id)
                line["cpu_baseline"] = {
                    "value": cp / dt, "unit": UNIT, "cores": 1, "kind": "port",
                    "sample": (f"{args.cpu_baseline_steps} full cycles of the same workload"
                               + ("" if stride_s == 1 else f" restricted to every {stride_s}-th velocity sample")
                               + f" ({ro.n_poses} poses each) through the oracle port, kd-tree rebuilt every cycle; "
                               f"index={'reference-vendored nanoflann 1.5.1 (oracle/_ref)' if use_ref else 'oracle bucket grid'}, glibc libm, "
                               f"1 thread; host has {os.cpu_count()} cores"),
                    "ms_per_cycle": 1e3 * dt / args.cpu_baseline_steps,
                    "stages_s_last_cycle": {"index_build": ora.timing[0], "rollout": ora.timing[1], "score": ora.timing[2]},
                    "best_id_matches_gpu": bool(stride_s != 1 or ro.best_id == r.best_id),
                }
        except Exception as exc:  # noqa: BLE001
            line["cpu_baseline"] = {"error": f"{type(exc).__name__}: {exc}"}

    if rank == 0:
        print(json.dumps(line), flush=True)
    lp.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
