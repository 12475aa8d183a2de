#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2h}
summ() {
  python - "$1" "$2" <<'PY'
import json, sys
W, F = sys.argv[1], sys.argv[2]
try:
    l=json.loads(open(F).read().strip().splitlines()[-1])
    s=l["stats_one_step"]; r=l["roofline"]; t=l["tree"]; nq=s["n_queries"]
    print("%s: value %.3f Mq/s e2e %.3f ms/step %.2f | seed %.2f select %.2f main %.2f | scanned/q %.0f pages/q %.0f evals/q %.0f ins/q %.0f | tc %d | parity %s/%s | clocks %s %s" % (
        W, l["value"]/1e6, l["e2e"]["value"]/1e6, l["ms_per_step"], s["ms_seed_scan"], s["ms_traverse"], s["ms_main_scan"], r["gpu_scanned_points_per_query"],
        s["pages_listed"]/nq, s["exact_reevaluations"]/nq, s["list_insertions"]/nq, s["prefilter_batches"], l.get("parity_mismatches"), l.get("parity_checked"),
        l["clocks"]["sm_mhz"], l["clocks"]["reasons"]))
except Exception as e:
    print(W, "no line:", e)
PY
}
run() {  # name workload env...
  local name=$1 W=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $W --steps 3 --warmup 3 --cpu-sample 64 > gpurun_out/bench_${W}_${name}_$TAG.json 2> gpurun_out/bench_${W}_${name}_$TAG.err
  summ "$W[$name]" gpurun_out/bench_${W}_${name}_$TAG.json; tail -2 gpurun_out/bench_${W}_${name}_$TAG.err
}
echo "== default"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_default_$TAG.json 2> gpurun_out/bench_default_$TAG.err; summ default gpurun_out/bench_default_$TAG.json; tail -2 gpurun_out/bench_default_$TAG.err
for W in gauss_10Mx32_q1M_K10 siftlike_1Mx128_q100k_K10 gauss_2Mx32_q200k_K100 gauss_2Mx32_q200k_K10; do
  for FP in 0 1 2; do run fp$FP $W SCHT_FIRST_PASS=$FP; done
done
run sleep500 gauss_10Mx32_q1M_K10 SCHT_FIRST_PASS=0 SCHT_RESCORER_SLEEP_NS=500
run sleep500 uniform_1Mx16_q100k_K10 SCHT_RESCORER_SLEEP_NS=500
echo "== pytest gpu (multi + parity)"
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q --timeout 200 --timeout-method thread -x > gpurun_out/pytest_gpu_$TAG.txt 2>&1; tail -5 gpurun_out/pytest_gpu_$TAG.txt
