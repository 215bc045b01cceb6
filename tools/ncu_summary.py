#!/usr/bin/env python
"""Summarise ncu output brought back in gpurun_out/ into small text files under profiles/ (tracked).

  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_launches.md
  python tools/ncu_summary.py kernel   gpurun_out/prof_plan.ncu-rep profiles/r01_plan_kernel.md
"""
import collections
import csv
import json
import subprocess
import sys

KEY_METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "l1tex__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__issue_inst0.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_active.avg",
]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        if row.get("Metric Name") == "gpu__time_duration.sum":
            agg[row["Kernel Name"].split("(")[0]].append(float(row["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write("| kernel | launches | avg us | total us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{k}` | {len(v)} | {sum(v)/len(v)/1e3:.2f} | {sum(v)/1e3:.1f} | {sum(v)/tot:.3f} |\n")
    print(open(dst).read())


def kernel(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = ["| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |", "|---|---|" + "---:|" * len(data)]
    name_col = hdr.index("Kernel Name")
    for m in KEY_METRICS:
        if m in hdr:
            i = hdr.index(m)
            out.append(f"| {m} | {units[i]} | " + " | ".join(r[i] for r in data) + " |")
    stalls = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((float(data[0][i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    with open(dst, "w") as f:
        f.write(f"kernel: `{data[0][name_col].split('(')[0]}`  (ncu --set full --clock-control none)\n\n")
        f.write("\n".join(out) + "\n\nwarp stall reasons (avg warps stalled per issue-active cycle, launch 0):\n\n")
        f.write("\n".join(f"- {n}: {v:.3f}" for v, n in stalls if v > 0.005) + "\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
