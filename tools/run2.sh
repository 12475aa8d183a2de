#!/bin/bash
# GPU pass: whole GPU test suite, default bench line, secondary workloads
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2b}
echo "== pytest gpu"
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 --timeout-method thread -x > gpurun_out/pytest_gpu_$TAG.txt 2>&1; tail -6 gpurun_out/pytest_gpu_$TAG.txt
grep -h "cfg3 K=" gpurun_out/pytest_gpu_$TAG.txt
echo "== bench default"
timeout 900 python bench.py > gpurun_out/bench_default_$TAG.json 2> gpurun_out/bench_default_$TAG.err; tail -1 gpurun_out/bench_default_$TAG.json | cut -c1-300; tail -3 gpurun_out/bench_default_$TAG.err
for W in gauss_2Mx32_q200k_K10 gauss_2Mx32_q200k_K100 gauss_2Mx32_q200k_K1 siftlike_1Mx128_q100k_K10 posture_like_78095x15_self_K10 gauss_10Mx32_q1M_K10; do
  echo "== bench $W"
  timeout 900 python bench.py --workload $W --steps 3 --warmup 3 > gpurun_out/bench_${W}_$TAG.json 2> gpurun_out/bench_${W}_$TAG.err
  python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/bench_${W}_$TAG.json").read().strip().splitlines()[-1])
    s=l["stats_one_step"]; r=l["roofline"]
    print("value %.3f Mq/s e2e %.3f Mq/s ms/step %.2f | seed %.2f select %.2f main %.2f | scanned/query gpu %.0f ref %s | tc %d | parity %s/%s | cpu %s q/s (%s) | clocks %s %s" % (
        l["value"]/1e6, l["e2e"]["value"]/1e6, l["ms_per_step"], s["ms_seed_scan"], s["ms_traverse"], s["ms_main_scan"], r["gpu_scanned_points_per_query"],
        r["ref_scanned_points_per_query"], s["prefilter_batches"], l.get("parity_mismatches"), l.get("parity_checked"), l.get("cpu_baseline",{}).get("value"), l.get("cpu_baseline",{}).get("kind"),
        l["clocks"]["sm_mhz"], l["clocks"]["reasons"]))
except Exception as e:
    print("no line:", e)
PY
  tail -2 gpurun_out/bench_${W}_$TAG.err
done
