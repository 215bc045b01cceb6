#!/usr/bin/env python
"""A few C2 cycles with one build of libb200lp.so (for ncu):  python tools/time_one.py tools/variants/lib_x.so [C2]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dddmr_navigation_b200 import LocalPlanner, make_query, synth
mk = {"C1": synth.c1_ramp, "C2": synth.c2_dense, "C3": synth.c3_multilevel}[sys.argv[2] if len(sys.argv) > 2 else "C2"]
sc = mk()
lp = LocalPlanner(sc.config, device=0, lib_path=os.path.abspath(sys.argv[1]))
lp.set_cloud(sc.cloud); lp.set_plan(sc.plan)
q = make_query(sc.pose, sc.twist)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for i in range(8):
    flush.zero_(); torch.cuda.synchronize()
    r = lp.plan(q)
print(os.path.basename(sys.argv[1]), r.best_id, lp.last_kernel_ms())
