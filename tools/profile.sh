#!/bin/bash
# ncu evidence of one bench step (GPU box only; the plain run must exit 0 first, B200_PROFILING.md).
# outputs: gpurun_out/launches_$TAG.csv, gpurun_out/prof_tc_$TAG.ncu-rep, gpurun_out/prof_seed_$TAG.ncu-rep
TAG=${1:-r01c}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
cd "$(dirname "$0")/.."
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_scan_tc -s 1 -c 1 -f -o gpurun_out/prof_tc_$TAG $CMD > gpurun_out/ncu_tc_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^k_scan$ -s 1 -c 1 -f -o gpurun_out/prof_seed_$TAG $CMD > gpurun_out/ncu_seed_$TAG.log 2>&1
ls -la gpurun_out/*$TAG*
