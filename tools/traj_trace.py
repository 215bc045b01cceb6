#!/usr/bin/env python
"""Schedule of plan_kernel at trajectory granularity (needs a -DB200LP_TRAJ_TRACE build, which overwrites cost / first_hit with
start / duration in ns):  python tools/traj_trace.py tools/variants_trace/lib_trace.so [C2|C4 [shard_rank shard_count]]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from dddmr_navigation_b200 import LocalPlanner, make_query, synth
name = sys.argv[2] if len(sys.argv) > 2 else "C2"
sc = synth.c3_multilevel(samples=(361.0, 361.0)) if name == "C4" else {"C1": synth.c1_ramp, "C2": synth.c2_dense, "C3": synth.c3_multilevel}[name]()
shard = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else None
lp = LocalPlanner(sc.config, device=0, lib_path=os.path.abspath(sys.argv[1]))
lp.set_cloud(sc.cloud); lp.set_plan(sc.plan)
q = make_query(sc.pose, sc.twist)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for i in range(6):
    flush.zero_(); torch.cuda.synchronize()
    r = lp.plan_shard(q, *shard) if shard else lp.plan(q)
km = lp.last_kernel_ms()
t = lp.read_trajectories()
if shard:
    _, b, e = lp.traj_count()
    t = {k: v[b:e] for k, v in t.items()}
start, dur, steps = t["cost"] / 1e3, t["first_hit_pose"].astype(np.float64) / 1e3, t["num_steps"]
hit = t["critic_scores"][:, 0]
end = start + dur
k0 = start.min()
print("kernel ms", km, "trajectories", len(start))
print(f"first trajectory starts {k0:.1f} us after the cycle's first CTA, last one ends at {end.max():.1f} us; cycle_ns {lp.last_cycle_ns()['cycle_ns'] / 1e3:.1f} us")
print(f"plan_kernel window from first start: {end.max() - k0:.1f} us; sum of durations {dur.sum():.0f} us = {dur.sum() / 2960:.1f} us per resident warp (2960)")
print("duration us: mean %.1f p50 %.1f p90 %.1f p99 %.1f max %.1f" % (dur.mean(), np.percentile(dur, 50), np.percentile(dur, 90), np.percentile(dur, 99), dur.max()))
for lo, hi in ((0, 25), (25, 50), (50, 75), (75, 90), (90, 100)):
    a, b = k0 + (end.max() - k0) * lo / 100, k0 + (end.max() - k0) * hi / 100
    m = (start >= a) & (start < b)
    running = ((start < (a + b) / 2) & (end > (a + b) / 2)).sum()
    print(f"  window {lo:3d}-{hi:3d}%: {m.sum():6d} started, mean dur {dur[m].mean() if m.any() else 0:6.1f} us, running at its middle: {running}")
late = np.argsort(end)[-10:]
print("last to finish: ", [(int(i), round(float(start[i] - k0), 1), round(float(dur[i]), 1), int(steps[i]), int(hit[i])) for i in late])
coll = hit >= 0
print(f"colliding {coll.sum()}: mean dur {dur[coll].mean():.1f}; free {(~coll).sum()}: mean dur {dur[~coll].mean():.1f}, max {dur[~coll].max():.1f}")
