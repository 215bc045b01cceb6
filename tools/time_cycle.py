#!/usr/bin/env python
"""Whole-cycle and per-kernel times (CUDA events) + the host clock around plan() of several builds (tools/variants/*.so):
   python tools/time_cycle.py [C2 C3 C1]"""
import glob, os, statistics, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dddmr_navigation_b200 import LocalPlanner, make_query, synth
MAKERS = {"C1": synth.c1_ramp, "C2": synth.c2_dense, "C3": synth.c3_multilevel}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
names = [a for a in sys.argv[1:] if a in MAKERS] or ["C2"]
for name in names:
    sc = MAKERS[name]()
    q = make_query(sc.pose, sc.twist)
    for lib in sorted(glob.glob(os.path.join(ROOT, "tools", "variants", "*.so"))):
        lp = LocalPlanner(sc.config, device=0, lib_path=lib)
        lp.set_cloud(sc.cloud); lp.set_plan(sc.plan)
        rows, wall = [], []
        for i in range(33):
            flush.zero_(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = lp.plan(q)
            w = 1e3 * (time.perf_counter() - t0)
            if i >= 3:
                km = lp.last_kernel_ms(); rows.append((lp.last_timing()["ms_plan_kernels"], km["prep_kernel"], km["cull_kernel"], km["plan_kernel"])); wall.append(w)
        med = [statistics.median(c) for c in zip(*rows)]
        print(f"{name} {os.path.basename(lib):22s} wall={statistics.median(wall):.4f} cycle={med[0]:.4f} prep={med[1]:.4f} cull={med[2]:.4f} plan={med[3]:.4f} best={r.best_id} coll={r.n_collided}", flush=True)
        lp.close()
