#!/usr/bin/env python
"""Host-clock time of b200lp_plan_batch (C5: 512 robots on the 8 M-point map) with the plan table in pageable and in pinned
host memory, next to the CUDA-event time of its kernels:  python tools/time_fleet_call.py"""
import os, statistics, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from dddmr_navigation_b200 import LocalPlanner, abi, make_query, synth
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
base = synth.c3_multilevel(samples=(20.0, 25.0))
c1 = synth.c1_ramp(n_points=1000)
poses, twists, plans, offs = synth.fleet_queries(512, region=(-28.0, 28.0, -20.0, 20.0), levels=(0.0, 3.0, 6.0), cloud=base.cloud)
qs = (abi.Query * 512)()
for i in range(512):
    qs[i] = make_query(poses[i], twists[i])
plans = np.ascontiguousarray(plans, np.float64); offs = np.ascontiguousarray(offs, np.int64)
pin = torch.from_numpy(plans).pin_memory(); pinned = pin.numpy()
lp = LocalPlanner(c1.config, device=0)
lp.set_cloud(base.cloud)
for rnd in range(2):
    for name, tab in (("pageable", plans), ("pinned", pinned)):
        wall, dev = [], []
        for i in range(15):
            flush.zero_(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = lp.plan_batch(qs, tab, offs)
            dt = time.perf_counter() - t0
            if i >= 3:
                wall.append(1e3 * dt); dev.append(lp.last_timing()["ms_plan_kernels"])
        print(f"C5 plan table {name:9s}: call {statistics.median(wall):.4f} ms, kernels {statistics.median(dev):.4f} ms, best0={res[0].best_id}", flush=True)
