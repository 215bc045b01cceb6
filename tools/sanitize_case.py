#!/usr/bin/env python
"""Small end-to-end exercise of every kernel (for `compute-sanitizer --tool memcheck python tools/sanitize_case.py`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dddmr_navigation_b200 import LocalPlanner, abi, make_query, synth
sc = synth.c1_ramp(n_points=20_000)
lp = LocalPlanner(sc.config)
lp.set_cloud(sc.cloud); lp.set_plan(sc.plan)
q = make_query(sc.pose, sc.twist)
r = lp.plan(q); print("plan", r.as_dict())
t = lp.read_trajectories(); p = lp.read_pose_batch(0, 8); print("poses", p["pose"].shape, lp.count_radius())
for rank in range(3):
    print("shard", rank, lp.plan_shard(q, rank, 3).best_id)
n = 6
poses, twists, plans, offs = synth.fleet_queries(n, region=(2.0, 20.0, -6.0, 6.0))
qs = (abi.Query * n)()
for i in range(n):
    qs[i] = make_query(poses[i], twists[i])
res = lp.plan_batch(qs, plans, offs); print("fleet", [x.best_id for x in res])
lp.set_global_plan(np.concatenate([sc.plan, sc.plan[-1:] + np.array([[0.05 * k, 0, 0, 0, 0, 0, 0] for k in range(1, 40)])]))
info = lp.prune_plan(sc.pose[:3], 3.0, 1.0); print("prune", info.as_dict(), lp.path_blocked(0.5).as_dict())
print("plan on device prune plan", lp.plan(q).best_id)
empty = LocalPlanner(sc.config); empty.set_plan(sc.plan); print("no cloud", empty.plan(q).best_id)
