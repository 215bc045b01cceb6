#!/bin/bash
# N-GPU bench lines of a round (one box, one rank per GPU): bash tools/evidence_n8.sh <tag> <N>
tag=${1:-x}; N=${2:-8}; out=gpurun_out; port=29541
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N "$@"; port=$((port+1)); }
run --no-cpu-baseline --latency-cycles 0 --observation-scans 0 > $out/bench_${tag}_n${N}_C2.json 2> $out/bench_${tag}_n${N}.err
run --workload C5 --no-cpu-baseline --latency-cycles 0 --observation-scans 0 > $out/bench_${tag}_n${N}_C5.json 2>> $out/bench_${tag}_n${N}.err
run --workload C4 --no-cpu-baseline --latency-cycles 0 --observation-scans 0 > $out/bench_${tag}_n${N}_C4.json 2>> $out/bench_${tag}_n${N}.err
for w in C2 C5 C4; do python -c "
import json,sys; d=json.load(open('$out/bench_${tag}_n${N}_$w.json')); print('$w', 'n', d['n_gpus'], 'value %.3e' % d['value'], 'ms', round(d['ms_per_step'],4), 'e2e %.3e' % d['e2e']['value'], d['scaling'], d['cpu_affinity'])" || tail -5 $out/bench_${tag}_n${N}.err; done
