#!/usr/bin/env python
"""profiles/traffic.json from the ncu summaries of one capture set:  python tools/make_traffic.py <tag> "<commit note>"
(reads profiles/<tag>_plan_kernel_{C1,C2,C3,C5}.txt as written by tools/ncu_summary.py; bench.py quotes the entries in
roofline.traffic / roofline.physical together with the commit they were captured at)."""
import json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, commit = sys.argv[1], sys.argv[2]
path = os.path.join(ROOT, "profiles", "traffic.json")
out = json.load(open(path))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for w in ("C1", "C2", "C3", "C5"):
    f = os.path.join(ROOT, "profiles", f"{tag}_plan_kernel_{w}.txt")
    if not os.path.exists(f):
        continue
    vals = {}
    for ln in open(f):
        m = re.match(r"\s+(\S+)\s+(\S+)\s+(\S+)\s*$", ln)
        if m:
            vals[m.group(1)] = (m.group(2), m.group(3))
    def num(k):
        u, v = vals[k]
        return float(v) * UNIT.get(u, 1)
    out[w] = {
        "plan_kernel_dram_bytes_per_launch": int(num("dram__bytes_read.sum") + num("dram__bytes_write.sum")),
        "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warps_eligible_per_cycle": num("smsp__warps_eligible.avg.per_cycle_active"),
        "warp_instructions": int(num("smsp__inst_executed.sum")),
        "kernel_us_under_ncu": num("gpu__time_duration.sum"),
        "source": f"profiles/{tag}_plan_kernel_{w}.txt (ncu --set full --clock-control none, one launch after an L2 flush)",
        "commit": commit,
    }
json.dump(out, open(path, "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k in ("C1", "C2", "C3", "C5")}, indent=1))
