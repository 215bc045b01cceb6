#!/usr/bin/env python
"""Time the plan cycle's kernels for several voxel-grid cell sizes (run on a B200):
    python tools/tune_grid.py [C1|C2|C3] ...
Prints one line per (workload, cell_xy, cell_z): prep / plan kernel ms (median of 10, L2 flushed), grid build ms."""
import dataclasses
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from dddmr_navigation_b200 import LocalPlanner, make_query, synth  # noqa: E402

MAKERS = {"C1": synth.c1_ramp, "C2": synth.c2_dense, "C3": synth.c3_multilevel}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name in (sys.argv[1:] or ["C2"]):
    sc = MAKERS[name]()
    q = make_query(sc.pose, sc.twist)
    ref = None
    for cxy, cz in [(0.25, 0.25), (0.25, 0.5), (0.25, 1.0), (0.2, 0.2), (0.2, 0.4), (0.15, 0.15), (0.15, 0.3), (0.125, 0.25), (0.1, 0.2), (0.1, 0.1), (0.35, 0.35), (0.5, 0.5)]:
        cfg = dataclasses.replace(sc.config, cell_xy=cxy, cell_z=cz)
        lp = LocalPlanner(cfg, device=0)
        lp.set_cloud(sc.cloud)
        lp.set_cloud(sc.cloud)
        gms = lp.last_timing()["ms_grid_build"]
        lp.set_plan(sc.plan)
        pk, pp = [], []
        for i in range(13):
            flush.zero_()
            torch.cuda.synchronize()
            r = lp.plan(q)
            if i >= 3:
                km = lp.last_kernel_ms()
                pk.append(km["plan_kernel"])
                pp.append(km["prep_kernel"])
        key = (r.best_id, r.best_cost, r.n_collided, r.n_poses)
        ref = ref or key
        print(f"{name} cell_xy={cxy} cell_z={cz} grid={lp.grid_info()['dims']} plan_kernel={statistics.median(pk):.4f} ms "
              f"prep={statistics.median(pp):.4f} ms grid_build={gms:.3f} ms poses={r.n_poses} same_result={key == ref}", flush=True)
        lp.close()
