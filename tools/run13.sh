#!/bin/bash
# GPU box: full validation of the final round-2 code (tests, smoke, every bench line)
cd "$(dirname "$0")/.."
bash tools/final_check.sh r02g
for W in gauss_2Mx32_q200k_K100 gauss_2Mx32_q200k_K1; do
  timeout 600 python bench.py --workload $W --steps 3 --warmup 3 > gpurun_out/bench_${W}_r02g.json 2> gpurun_out/bench_${W}_r02g.err; tail -1 gpurun_out/bench_${W}_r02g.json | cut -c1-160
done
