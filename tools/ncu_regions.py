#!/usr/bin/env python
"""Per source line (and per enclosing function) warp instructions and stall samples of one kernel of an ncu report.
    ncu -i X.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all lib.so; nvdisasm -g -c *.cubin > dis.txt
    python tools/ncu_regions.py src.csv dis.txt plan_kernel [top_n]
Functions are found by scanning csrc/*.cuh for `__device__` / `__global__` definitions (the last one starting at or before a line)."""
import collections, csv, glob, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_lines import load_lines

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def functions():
    out = {}
    for f in glob.glob(os.path.join(ROOT, "dddmr_navigation_b200", "csrc", "*")):
        if not f.endswith((".cuh", ".h", ".cu")):
            continue
        starts = []
        for i, ln in enumerate(open(f), 1):
            m = re.match(r"\s*(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:__host__\s+)?(?:__device__|__global__)[^;(]*?\b(\w+)\s*\(", ln)
            if m and not ln.strip().startswith("//"):
                starts.append((i, m.group(1)))
            m2 = re.match(r"\s*(?://\s*----\s*(.+?)\s*-+\s*$)", ln)
            if m2 and "plan_kernel" in open(f).read()[:0] + "":
                pass
        out[os.path.basename(f)] = starts
    return out


def main():
    src, dis, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    off2line = load_lines(dis, kernel)
    rows = list(csv.reader(open(src)))
    sect = [i for i, r in enumerate(rows) if r and r[0] == "Address" and i > 0 and kernel in rows[i - 1][1]]
    a = sect[0]
    hdr = rows[a]
    ix = {h: i for i, h in enumerate(hdr)}
    body = []
    for r in rows[a + 1:]:
        if len(r) != len(hdr) or r[0] in ("Address", "Kernel Name"):
            break
        body.append(r)
    base = int(body[0][0], 16)
    fn = functions()
    inst, samp = collections.Counter(), collections.Counter()
    finst, fsamp, fsass = collections.Counter(), collections.Counter(), collections.Counter()
    for r in body:
        key = off2line.get(int(r[0], 16) - base, ("?", 0))
        e, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
        inst[key] += e
        samp[key] += s
        name = key[0]
        for ln, f in fn.get(key[0], []):
            if ln <= key[1]:
                name = f"{key[0]}:{f}"
        finst[name] += e
        fsamp[name] += s
        fsass[name] += 1
    ti, ts = sum(inst.values()), sum(samp.values())
    print(f"{kernel}: {len(body)} SASS instructions, {ti} warp instructions executed, {ts} samples")
    print("-- by function")
    for k, v in finst.most_common():
        print(f"{k:48s} inst {v:>10d} {100 * v / ti:5.1f}%  samples {100 * fsamp[k] / max(ts, 1):5.1f}%  sass {fsass[k]}")
    print("-- by line (top samples)")
    for key, v in samp.most_common(top):
        print(f"{key[0]}:{key[1]:<5d} samples {v:>6d} {100 * v / max(ts, 1):5.1f}%   inst {inst[key]:>10d} {100 * inst[key] / ti:5.1f}%")


if __name__ == "__main__":
    main()
