#!/usr/bin/env python
"""set_cloud timing (pinned PointXYZI upload + grid build) for the builds under tools/variants/: python tools/time_setcloud.py [C2 C3]"""
import glob, os, statistics, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from dddmr_navigation_b200 import LocalPlanner, synth
MAKERS = {"C1": synth.c1_ramp, "C2": synth.c2_dense, "C3": synth.c3_multilevel}
for name in (sys.argv[1:] or ["C2"]):
    sc = MAKERS[name]()
    t = torch.from_numpy(np.ascontiguousarray(sc.cloud)).pin_memory(); cloud = t.numpy()
    for lib in sorted(glob.glob(os.path.join(ROOT, "tools", "variants", "*.so"))):
        lp = LocalPlanner(sc.config, device=0, lib_path=lib)
        wall, up, gr = [], [], []
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        for i in range(15):
            if os.environ.get("FLUSH"):
                flush.zero_(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            lp.set_cloud_ptr(cloud.ctypes.data, cloud.shape[0], cloud.shape[1] * 4)
            dt = time.perf_counter() - t0
            if i >= 3:
                tm = lp.last_timing(); wall.append(1e3 * dt); up.append(tm["ms_upload"]); gr.append(tm["ms_grid_build"])
        print(f"{name} {os.path.basename(lib):18s} set_cloud wall {statistics.median(wall):.4f} ms  upload {statistics.median(up):.4f}  grid {statistics.median(gr):.4f}", flush=True)
        lp.close()
