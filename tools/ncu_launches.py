#!/usr/bin/env python
"""ncu launch-list CSV (`--metrics gpu__time_duration.sum --csv --log-file X.csv`) -> markdown table per kernel.
usage: ncu_launches.py launches.csv > profiles/rNN_launches.md"""
import collections, csv, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
    v = float(r[ix["Metric Value"]]) / (1e3 if r[ix["Metric Unit"]] == "ns" else 1.0)
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
print("| kernel | launches | avg us | total us | share |\n|---|---:|---:|---:|---:|")
for k, v in tot.most_common():
    print(f"| `{k}` | {cnt[k]} | {v / cnt[k]:.2f} | {v:.1f} | {v / T:.3f} |")
