#!/usr/bin/env python
"""Phase timestamps of prep_kernel (needs tools/variants_trace/lib_prep_trace.so built with -DB200LP_PREP_TRACE):
   python tools/prep_trace.py [C2|C4] [shard_rank shard_count]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dddmr_navigation_b200 import LocalPlanner, make_query, synth
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
sc = synth.c2_dense() if name == "C2" else synth.c3_multilevel(n_points=2_000_000, samples=(361.0, 361.0))
lp = LocalPlanner(sc.config, device=0, lib_path=os.path.join(ROOT, "tools", "variants_trace", "lib_prep_trace.so"))
lp.set_cloud(sc.cloud); lp.set_plan(sc.plan)
q = make_query(sc.pose, sc.twist)
for i in range(3):
    print(f"--- {name} cycle {i}", flush=True)
    if len(sys.argv) > 3:
        lp.plan_shard(q, int(sys.argv[2]), int(sys.argv[3]))
    else:
        lp.plan(q)
    print(lp.last_kernel_ms(), flush=True)
lp.close()
