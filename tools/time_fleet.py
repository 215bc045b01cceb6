#!/usr/bin/env python
"""Kernel times of a fleet cycle (C5: 512 robots on the 8 M-point map) and of the C4 sample set for several builds of
libb200lp.so (tools/variants/*.so):  python tools/time_fleet.py"""
import glob, os, statistics, sys
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
c4 = synth.c3_multilevel(samples=(361.0, 361.0))
for lib in sorted(glob.glob(os.path.join(ROOT, "tools", "variants", "*.so"))):
    for name, cfg in (("C5", c1.config), ("C4", c4.config)):
        lp = LocalPlanner(cfg, device=0, lib_path=lib)
        lp.set_cloud(base.cloud)
        if name == "C4":
            lp.set_plan(c4.plan)
        tot, pk, pp = [], [], []
        for i in range(13):
            flush.zero_(); torch.cuda.synchronize()
            if name == "C5":
                res = lp.plan_batch(qs, plans, offs); r = res[0]
            else:
                r = lp.plan(make_query(c4.pose, c4.twist))
            if i >= 3:
                km = lp.last_kernel_ms(); tot.append(lp.last_timing()["ms_plan_kernels"]); pk.append(km["plan_kernel"]); pp.append(km["prep_kernel"])
        print(f"{name} {os.path.basename(lib):22s} cycle={statistics.median(tot):.4f} ms plan_kernel={statistics.median(pk):.4f} prep={statistics.median(pp):.4f} best0={r.best_id}", flush=True)
        lp.close()
