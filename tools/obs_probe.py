import sys, time, statistics, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from dddmr_navigation_b200 import LocalPlanner, synth
sc = synth.c2_dense()
scan, b2s, g2b = synth.lidar_scan(n_beams=128, n_azimuth=2048)
t = torch.from_numpy(scan).pin_memory(); hscan = t.numpy()
tc = torch.from_numpy(np.ascontiguousarray(sc.cloud)).pin_memory(); cloud = tc.numpy()
lp = LocalPlanner(sc.config, device=0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
FLUSH = len(sys.argv) > 1
def run(tag):
    for _ in range(10): lp.sensor_observation(0, hscan, b2s, g2b, 10.0, 2.0)
    w, d = [], []
    for _ in range(50):
        if FLUSH: flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter(); oi = lp.sensor_observation(0, hscan, b2s, g2b, 10.0, 2.0); w.append(1e3*(time.perf_counter()-t0)); d.append(oi.ms_device)
    print(tag, "wall p50 %.3f device p50 %.3f" % (statistics.median(w), statistics.median(d)), flush=True)
dscan = torch.empty_like(t, device="cuda")
def h2d():
    w = []
    for _ in range(30):
        if FLUSH: flush.zero_()
        torch.cuda.synchronize(); t0 = time.perf_counter(); dscan.copy_(t, non_blocking=True); torch.cuda.synchronize(); w.append(1e3*(time.perf_counter()-t0))
    print("plain 4 MB pinned H2D copy p50 %.3f ms" % statistics.median(w), flush=True)
h2d()
run("fresh ctx, no cloud yet      ")
lp.set_cloud_ptr(cloud.ctypes.data, cloud.shape[0], 32); lp.last_timing()
run("after one packed set_cloud   ")
for _ in range(20): lp.set_cloud_ptr(cloud.ctypes.data, cloud.shape[0], 32)
lp.last_timing()
run("after 20 more                ")
time.sleep(0.5)
run("after 0.5 s sleep            ")
