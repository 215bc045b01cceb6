#!/usr/bin/env python
"""A few cycles of one BASELINE workload with the in-tree library (for ncu):  python tools/run_workload.py C1|C2|C3|C4|C5 [cycles]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from dddmr_navigation_b200 import LocalPlanner, abi, make_query, synth
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
cycles = int(sys.argv[2]) if len(sys.argv) > 2 else 6
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
if name == "C5":
    base = synth.c3_multilevel(samples=(20.0, 25.0))
    c1 = synth.c1_ramp(n_points=1000)
    poses, twists, plans, offs = synth.fleet_queries(512, region=(-28.0, 28.0, -20.0, 20.0), levels=(0.0, 3.0, 6.0), cloud=base.cloud)
    qs = (abi.Query * 512)()
    for i in range(512):
        qs[i] = make_query(poses[i], twists[i])
    plans = np.ascontiguousarray(plans, np.float64); offs = np.ascontiguousarray(offs, np.int64)
    lp = LocalPlanner(c1.config, device=0)
    lp.set_cloud(base.cloud)
    fn = lambda: lp.plan_batch(qs, plans, offs)[0]
else:
    sc = {"C1": synth.c1_ramp, "C2": synth.c2_dense, "C3": synth.c3_multilevel,
          "C4": lambda: synth.c3_multilevel(samples=(361.0, 361.0))}[name]()
    lp = LocalPlanner(sc.config, device=0)
    lp.set_cloud(sc.cloud); lp.set_plan(sc.plan)
    q = make_query(sc.pose, sc.twist)
    fn = lambda: lp.plan(q)
for i in range(cycles):
    flush.zero_(); torch.cuda.synchronize()
    r = fn()
print(name, r.best_id, r.n_poses, lp.last_kernel_ms())
