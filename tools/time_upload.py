#!/usr/bin/env python
"""Time set_cloud (host call until return, and until the grid is ready) of the builds under tools/variants for several
B200LP_PACK_THREADS values:  python tools/time_upload.py [C2|C3] [threads ...]"""
import glob, os, statistics, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from dddmr_navigation_b200 import LocalPlanner, synth
MAKERS = {"C1": synth.c1_ramp, "C2": synth.c2_dense, "C3": synth.c3_multilevel}
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
threads = [int(a) for a in sys.argv[2:]] or [8]
sc = MAKERS[name]()
t = torch.from_numpy(np.ascontiguousarray(sc.cloud)).pin_memory(); cloud = t.numpy()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
libs = sorted(glob.glob(os.path.join(ROOT, "tools", "variants", "*.so")))
for rnd in range(2):
    for T in threads:
        os.environ["B200LP_PACK_THREADS"] = str(T)
        for lib in libs:
            lp = LocalPlanner(sc.config, device=0, lib_path=lib)
            ret, ready, up, gb = [], [], [], []
            for i in range(16):
                flush.zero_(); torch.cuda.synchronize()
                t0 = time.perf_counter()
                lp.set_cloud_ptr(cloud.ctypes.data, cloud.shape[0], cloud.shape[1] * 4)
                t1 = time.perf_counter()
                tm = lp.last_timing()
                t2 = time.perf_counter()
                if i >= 4:
                    ret.append(1e3 * (t1 - t0)); ready.append(1e3 * (t2 - t0)); up.append(tm["ms_upload"]); gb.append(tm["ms_grid_build"])
            print(f"{name} r{rnd} T={T:2d} {os.path.basename(lib):16s} returns {statistics.median(ret):.3f}  grid ready {statistics.median(ready):.3f}  "
                  f"(device: upload {statistics.median(up):.3f} + grid {statistics.median(gb):.3f})", flush=True)
            lp.close()
