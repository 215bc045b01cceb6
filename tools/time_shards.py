#!/usr/bin/env python
"""Per-kernel times of every shard of a sample-sharded C4 cycle on ONE GPU (what each of W ranks would run):
   python tools/time_shards.py [W] [cut shares ...]"""
import os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dddmr_navigation_b200 import LocalPlanner, make_query, synth
W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cuts = [float(v) for v in sys.argv[2:]]
sc = synth.c3_multilevel(samples=(361.0, 361.0))
lp = LocalPlanner(sc.config, device=0)
lp.set_cloud(sc.cloud); lp.set_plan(sc.plan)
if cuts:
    lp.set_shard_cuts(cuts)
q = make_query(sc.pose, sc.twist)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(fn):
    tot, pk, pp, ns, cu = [], [], [], [], []
    for i in range(13):
        flush.zero_(); torch.cuda.synchronize()
        r = fn()
        if i >= 3:
            km = lp.last_kernel_ms(); tot.append(lp.last_timing()["ms_plan_kernels"]); pk.append(km["plan_kernel"]); pp.append(km["prep_kernel"]); cu.append(km["cull_kernel"])
            ns.append(lp.last_cycle_ns()["cycle_ns"] / 1e6)
    return r, statistics.median(tot), statistics.median(pp), statistics.median(pk), statistics.median(ns), statistics.median(cu)
r, t, pp, pk, ns, cu = run(lambda: lp.plan(q))
print(f"whole: cycle={t:.4f} prep={pp:.4f} cull={cu:.4f} plan={pk:.4f} device_ns={ns:.4f} traj={r.n_traj} poses={r.n_poses}")
for k in range(W):
    r, t, pp, pk, ns, cu = run(lambda: lp.plan_shard(q, k, W))
    print(f"shard {k}/{W}: cycle={t:.4f} prep={pp:.4f} cull={cu:.4f} plan={pk:.4f} device_ns={ns:.4f} traj={r.n_traj} poses={r.n_poses} ranges={lp.traj_count()}")
