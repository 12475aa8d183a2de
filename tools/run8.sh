#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2j}
summ() {
  python - "$1" "$2" <<'PY'
import json, sys
W, F = sys.argv[1], sys.argv[2]
try:
    l=json.loads(open(F).read().strip().splitlines()[-1])
    s=l["stats_one_step"]; r=l["roofline"]; nq=s["n_queries"]
    print("%s: value %.3f Mq/s ms/step %.2f | seed %.2f select %.2f main %.2f | pages/q %.0f evals/q %.0f ins/q %.0f launches %d" % (
        W, l["value"]/1e6, l["ms_per_step"], s["ms_seed_scan"], s["ms_traverse"], s["ms_main_scan"], s["pages_listed"]/nq, s["exact_reevaluations"]/nq, s["list_insertions"]/nq, s["kernel_launches"]))
except Exception as e:
    print(W, "no line:", e)
PY
}
run() {
  local name=$1 W=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${W}_${name}_$TAG.json 2> gpurun_out/bench_${W}_${name}_$TAG.err
  summ "$W[$name]" gpurun_out/bench_${W}_${name}_$TAG.json; tail -2 gpurun_out/bench_${W}_${name}_$TAG.err
}
for W in gauss_10Mx32_q1M_K10 gauss_2Mx32_q200k_K10; do
  run split1 $W SCHT_TC_SPLIT=1
  run split4 $W SCHT_TC_SPLIT=4
  run split8 $W SCHT_TC_SPLIT=8
done
