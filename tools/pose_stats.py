#!/usr/bin/env python
"""Distribution of |radiusSearch(pose,1.0)| over the poses of a workload (run on a B200)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from dddmr_navigation_b200 import LocalPlanner, make_query, synth
MAKERS = {"C1": synth.c1_ramp, "C2": synth.c2_dense, "C3": synth.c3_multilevel}
for name in (sys.argv[1:] or ["C2"]):
    sc = MAKERS[name]()
    lp = LocalPlanner(sc.config, device=0)
    lp.set_cloud(sc.cloud); lp.set_plan(sc.plan)
    r = lp.plan(make_query(sc.pose, sc.twist))
    t = lp.read_trajectories()
    d = lp.read_pose_batch(0, r.n_traj, 0, fields=["n_r1", "collide"])
    nr1, col = d["n_r1"], d["collide"]
    fh = t["first_hit_pose"]; steps = t["num_steps"]
    scored = np.where(fh >= 0, fh + 1, steps).sum()
    print(name, "poses", nr1.size, "n_r1==0: %.3f" % (nr1 == 0).mean(), "n_r1<32: %.3f" % (nr1 < 32).mean(), "mean", nr1.mean(), "median", np.median(nr1),
          "p90", np.percentile(nr1, 90), "collide frac %.3f" % col.mean(), "poses up to first hit: %d (%.3f)" % (scored, scored / nr1.size))
    lp.close()
