#!/usr/bin/env python
"""Time plan/prep kernels of several builds of libb200lp.so (tools/variants/*.so) on C2 / C1:  python tools/time_variants.py [C2 C1]"""
import glob, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dddmr_navigation_b200 import LocalPlanner, make_query, synth
MAKERS = {"C1": synth.c1_ramp, "C2": synth.c2_dense, "C3": synth.c3_multilevel,
          "C4": lambda: synth.c3_multilevel(samples=(361.0, 361.0))}  # "C4" also times shards 0, 4 and 7 of 8
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
names = [a for a in sys.argv[1:] if a in MAKERS] or ["C2"]
for name in names:
    sc = MAKERS[name]()
    q = make_query(sc.pose, sc.twist)
    for lib in sorted(glob.glob(os.path.join(ROOT, "tools", "variants", "*.so"))):
        lp = LocalPlanner(sc.config, device=0, lib_path=lib)
        lp.set_cloud(sc.cloud); lp.set_plan(sc.plan)
        pk, pp, cu = [], [], []
        for i in range(23):
            flush.zero_(); torch.cuda.synchronize()
            r = lp.plan(q)
            if i >= 3:
                km = lp.last_kernel_ms(); pk.append(km["plan_kernel"]); pp.append(km["prep_kernel"]); cu.append(km["cull_kernel"])
        print(f"{name} {os.path.basename(lib):20s} plan_kernel={statistics.median(pk):.4f} ms (min {min(pk):.4f}) prep={statistics.median(pp):.4f} cull={statistics.median(cu):.4f} ms best={r.best_id} cost={r.best_cost:.9f} coll={r.n_collided}", flush=True)
        for k in ((0, 4, 7) if name == "C4" else ()):
            pk, cy = [], []
            for i in range(23):
                flush.zero_(); torch.cuda.synchronize()
                r = lp.plan_shard(q, k, 8)
                if i >= 3:
                    pk.append(lp.last_kernel_ms()["plan_kernel"]); cy.append(lp.last_cycle_ns()["cycle_ns"] / 1e6)
            print(f"   shard {k}/8: plan_kernel={statistics.median(pk):.4f} ms, device cycle {statistics.median(cy):.4f} ms", flush=True)
        lp.close()
