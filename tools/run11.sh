#!/bin/bash
# GPU box: ncu evidence of the final round-2 code (cfg2 and Gaussian clusters 2Mx32), then the default bench line with the drop-in record
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash tools/profile.sh r02b_cfg2 uniform_1Mx16_q100k_K10 > gpurun_out/profile_r02b_cfg2.log 2>&1; tail -3 gpurun_out/profile_r02b_cfg2.log
bash tools/profile.sh r02b_gauss2M gauss_2Mx32_q200k_K10 k_select_pages > gpurun_out/profile_r02b_gauss2M.log 2>&1; tail -3 gpurun_out/profile_r02b_gauss2M.log
python bench.py > gpurun_out/bench_default_r02b.json 2> gpurun_out/bench_default_r02b.err; tail -1 gpurun_out/bench_default_r02b.json | cut -c1-200; tail -3 gpurun_out/bench_default_r02b.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/bench_default_r02b.json").read().strip().splitlines()[-1])
print("e2e_dropin:", l.get("e2e_dropin"))
PY
