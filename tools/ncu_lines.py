#!/usr/bin/env python
"""Join an ncu source-page CSV (SASS view) with nvdisasm -g line info: per source line, warp instructions
executed and stall samples. Usage:
    ncu -i X.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all libb200lp.so; nvdisasm -g -c b200lp.sm_100a.cubin > dis.txt
    python tools/ncu_lines.py src.csv dis.txt plan_kernel [top_n]
"""
import collections
import csv
import re
import sys


def load_lines(dis, kernel):
    off2line = {}
    cur = None
    inside = False
    for ln in open(dis):
        if ln.startswith(".text."):
            inside = kernel in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m:
            off2line[int(m.group(1), 16)] = cur
    return off2line


def main():
    src, dis, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    off2line = load_lines(dis, kernel)
    rows = list(csv.reader(open(src)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    first, seen = [], set()
    for r in rows[2:]:
        if len(r) != len(hdr) or r[0] == "Address" or r[0] in seen:
            break
        seen.add(r[0])
        first.append(r)
    base = int(first[0][0], 16)
    inst, samp = collections.Counter(), collections.Counter()
    for r in first:
        key = off2line.get(int(r[0], 16) - base, ("?", 0))
        inst[key] += int(r[ix["Instructions Executed"]])
        samp[key] += int(r[ix["# Samples"]])
    ti, ts = sum(inst.values()), sum(samp.values())
    print(f"total warp instructions {ti}, samples {ts}")
    for key, v in samp.most_common(top):
        print(f"{key[0]}:{key[1]:<5d} samples {v:>6d} {100 * v / ts:5.1f}%   inst {inst[key]:>10d} {100 * inst[key] / ti:5.1f}%")


if __name__ == "__main__":
    main()
