#!/bin/bash
# ncu evidence of one HEAD on one GPU: launch list of the default bench + full captures of plan_kernel on C1/C2/C3/C5 and of the
# grid / prep kernels on C2.  usage (on the GPU box): bash tools/evidence_r02.sh <tag>
tag=${1:-x}; out=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/launches_$tag.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --latency-cycles 0 --observation-scans 3 --no-multi --no-physical > $out/ncu_$tag.log 2>&1
for w in C2 C1 C3 C5; do
  ncu --set full --clock-control none --import-source on -k regex:'plan_kernel' --launch-skip 4 -c 1 -o $out/prof_${tag}_$w -f python tools/run_workload.py $w 6 > $out/ncu_${tag}_$w.log 2>&1
  ls -la $out/prof_${tag}_$w.ncu-rep
done
ncu --set full --clock-control none --import-source on -k regex:'prep_kernel|hist_kernel|scatter_kernel|sat_y_kernel|sat_z_kernel|scan_' --launch-skip 0 -c 12 -o $out/prof_${tag}_grid -f python tools/run_workload.py C2 2 > $out/ncu_${tag}_grid.log 2>&1
ls -la $out/prof_${tag}_grid.ncu-rep
