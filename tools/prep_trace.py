#!/usr/bin/env python
"""Phase timestamps of prep_kernel (needs tools/variants/lib_trace.so built with -DB200LP_PREP_TRACE)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dddmr_navigation_b200 import LocalPlanner, make_query, synth
for mk in (synth.c2_dense,):
    sc = mk()
    lp = LocalPlanner(sc.config, device=0, lib_path=os.path.join(ROOT, "tools", "variants_trace", "lib_trace.so"))
    lp.set_cloud(sc.cloud); lp.set_plan(sc.plan)
    q = make_query(sc.pose, sc.twist)
    for i in range(3):
        print(f"--- {sc.name} cycle {i}", flush=True)
        lp.plan(q)
        print(lp.last_kernel_ms(), flush=True)
    lp.close()
