#!/bin/bash
# ncu evidence of one bench step (GPU box only; the plain run must exit 0 first, B200_PROFILING.md).
# usage: tools/profile.sh TAG [WORKLOAD [extra kernel regex]]
# outputs: gpurun_out/launches_$TAG.csv, gpurun_out/prof_tc_$TAG.ncu-rep, gpurun_out/prof_seed_$TAG.ncu-rep (+ prof_x_$TAG for the extra kernel)
TAG=${1:-r02a}
W=${2:-uniform_1Mx16_q100k_K10}
X=$3
CMD="python bench.py --workload $W --steps 1 --warmup 1 --no-cpu-baseline"
cd "$(dirname "$0")/.."
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
# the LAST k_scan_tc launch of the command is the main scan of the instrumented step: skip the earlier ones
NTC=$(grep -c "k_scan_tc" gpurun_out/launches_$TAG.csv); NSEED=$(grep -c '"k_scan<\|k_scan(' gpurun_out/launches_$TAG.csv)
ncu --set full --clock-control none --import-source on -k regex:k_scan_tc -s $((NTC > 0 ? NTC - 1 : 0)) -c 1 -f -o gpurun_out/prof_tc_$TAG $CMD > gpurun_out/ncu_tc_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^k_scan$ -s 1 -c 1 -f -o gpurun_out/prof_seed_$TAG $CMD > gpurun_out/ncu_seed_$TAG.log 2>&1
if [ -n "$X" ]; then
  NX=$(grep -c "$X" gpurun_out/launches_$TAG.csv)
  ncu --set full --clock-control none --import-source on -k regex:$X -s $((NX > 0 ? NX - 1 : 0)) -c 1 -f -o gpurun_out/prof_x_$TAG $CMD > gpurun_out/ncu_x_$TAG.log 2>&1
fi
ls -la gpurun_out/*$TAG*
