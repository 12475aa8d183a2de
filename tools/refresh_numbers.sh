#!/bin/bash
# GPU box only: ncu evidence of the default bench step (tools/profile.sh TAG) and one bench line per secondary workload.
TAG=${1:-r01k}
cd "$(dirname "$0")/.."
bash tools/profile.sh $TAG > gpurun_out/profile_$TAG.log 2>&1
tail -3 gpurun_out/profile_$TAG.log
for w in posture_like_78095x15_self_K10 siftlike_1Mx128_q100k_K10 gauss_2Mx32_q200k_K1 gauss_2Mx32_q200k_K10 gauss_2Mx32_q200k_K100; do
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/err_$w.txt
  python - gpurun_out/bench_$w.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); s = d["stats_one_step"]
print(sys.argv[1], round(d["value"]), round(d["ms_per_step"], 2), {k: round(s[k], 2) for k in ("ms_seed_scan", "ms_traverse", "ms_main_scan")}, d["clocks"]["sm_mhz"])
PY
done
