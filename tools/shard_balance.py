#!/usr/bin/env python
"""Kernel time of every sample shard of the C4 workload, one after the other on one GPU: how balanced is the split?"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dddmr_navigation_b200 import LocalPlanner, make_query, synth
sc = synth.c3_multilevel(samples=(361.0, 361.0))
lp = LocalPlanner(sc.config, device=0)
lp.set_cloud(sc.cloud); lp.set_plan(sc.plan)
q = make_query(sc.pose, sc.twist)
for W in (int(a) for a in (sys.argv[1:] or ["8"])):
    rows = []
    for rank in range(W):
        ms = []
        for i in range(6):
            r = lp.plan_shard(q, rank, W)
            if i >= 2:
                ms.append(lp.last_timing()["ms_plan_kernels"])
        rows.append((rank, r.n_traj, r.n_poses, statistics.median(ms)))
    tot = sum(r[3] for r in rows)
    print(f"W={W}: max {max(r[3] for r in rows):.3f} ms, mean {tot / W:.3f} ms, imbalance {max(r[3] for r in rows) / (tot / W):.2f}x")
    for r in rows:
        print("   rank %d: %6d trajectories %8d poses %.3f ms" % r)
