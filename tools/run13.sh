#!/bin/bash
# GPU box: full validation of the final round-2 code + ncu captures of the clustered workload with the final kernels
cd "$(dirname "$0")/.."
bash tools/final_check.sh r02f
bash tools/profile.sh r02c_gauss2M gauss_2Mx32_q200k_K10 k_select_pages > gpurun_out/profile_r02c_gauss2M.log 2>&1; tail -2 gpurun_out/profile_r02c_gauss2M.log
for W in gauss_2Mx32_q200k_K100 gauss_2Mx32_q200k_K1; do
  timeout 600 python bench.py --workload $W --steps 3 --warmup 3 > gpurun_out/bench_${W}_r02f.json 2> gpurun_out/bench_${W}_r02f.err; tail -1 gpurun_out/bench_${W}_r02f.json | cut -c1-160
done
