#!/usr/bin/env python
"""A/B/A/B of set_cloud for the builds under tools/variants (order effects are large: link power states, first touch)."""
import glob, os, statistics, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from dddmr_navigation_b200 import LocalPlanner, synth
MAKERS = {"C1": synth.c1_ramp, "C2": synth.c2_dense, "C3": synth.c3_multilevel}
name = sys.argv[1] if len(sys.argv) > 1 else "C1"
sc = MAKERS[name]()
t = torch.from_numpy(np.ascontiguousarray(sc.cloud)).pin_memory(); cloud = t.numpy()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
libs = sorted(glob.glob(os.path.join(ROOT, "tools", "variants", "*.so")))
lps = [LocalPlanner(sc.config, device=0, lib_path=l) for l in libs]
for rnd in range(3):
    for lib, lp in zip(libs, lps):
        up, wall = [], []
        for i in range(12):
            flush.zero_(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            lp.set_cloud_ptr(cloud.ctypes.data, cloud.shape[0], cloud.shape[1] * 4)
            tm = lp.last_timing()
            wall.append(1e3 * (time.perf_counter() - t0)); up.append(tm["ms_upload"])
        print(f"{name} round {rnd} {os.path.basename(lib):14s} upload {statistics.median(up):.4f} wall(incl. grid wait) {statistics.median(wall):.4f}", flush=True)
