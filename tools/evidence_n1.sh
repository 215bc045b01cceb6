#!/bin/bash
# One-GPU evidence set of a round: GPU tests, smoke, bench lines for every workload + the reference arm, ncu launch list and a
# full capture of the cycle + grid + observation kernels. usage (on the GPU box): bash tools/evidence_n1.sh <tag>
tag=${1:-x}
out=gpurun_out
python -m pytest tests -m gpu -x -q > $out/gpu_tests_$tag.log 2>&1; tail -3 $out/gpu_tests_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; tail -2 $out/smoke_$tag.log
python bench.py > $out/bench_${tag}.json 2> $out/bench_${tag}.err || tail -5 $out/bench_${tag}.err
for w in C1 C3 C4 C5; do python bench.py --workload $w --no-cpu-baseline --latency-cycles 0 --observation-scans 0 > $out/bench_${tag}_$w.json 2>> $out/bench_${tag}.err; done
python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_${tag}_ref.json 2>> $out/bench_${tag}.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_$tag.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --latency-cycles 0 --observation-scans 3 > $out/ncu_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'prep_kernel|plan_kernel|hist_kernel|scatter_kernel|obs_scatter_kernel|obs_key_kernel' --launch-skip 40 -c 8 -o $out/prof_$tag -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline --latency-cycles 0 --observation-scans 3 > $out/ncu_${tag}2.log 2>&1
ls -la $out/prof_$tag.ncu-rep
