import os, sys, time
sys.path.insert(0, "/root/repo")
import torch
from dddmr_navigation_b200 import LocalPlanner, make_query, synth
for name in ("C2", "C4"):
    sc = synth.c2_dense() if name == "C2" else synth.c3_multilevel(samples=(361.0, 361.0))
    lp = LocalPlanner(sc.config, device=0)
    lp.set_cloud(sc.cloud); lp.set_plan(sc.plan)
    q = make_query(sc.pose, sc.twist)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for i in range(8):
        flush.zero_(); torch.cuda.synchronize()
        t0 = time.perf_counter(); lp.plan(q); t1 = time.perf_counter()
        print(f"{name} python call {1e6*(t1-t0):.1f} us", file=sys.stderr)
    if name == "C4":
        for i in range(6):
            flush.zero_(); torch.cuda.synchronize()
            t0 = time.perf_counter(); lp.plan_shard(q, 4, 8); t1 = time.perf_counter()
            print(f"C4 shard 4/8 python call {1e6*(t1-t0):.1f} us", file=sys.stderr)
    lp.close()
