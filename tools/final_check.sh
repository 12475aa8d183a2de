#!/bin/bash
# GPU box only: what the driver runs at round end -- GPU tests, smoke(), the bench line (with the CPU baseline) and the reference arm --
# plus the bench lines of the other BASELINE workloads.  usage: tools/final_check.sh TAG
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02}
timeout 2400 python -m pytest tests -m gpu -x -q --timeout 900 --timeout-method thread > gpurun_out/pytest_gpu_final_$TAG.txt 2>&1; tail -3 gpurun_out/pytest_gpu_final_$TAG.txt; grep -h "cfg3 K=" gpurun_out/pytest_gpu_final_$TAG.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_default_$TAG.json 2> gpurun_out/bench_default_$TAG.err; tail -1 gpurun_out/bench_default_$TAG.json | cut -c1-330
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2> gpurun_out/bench_reference_$TAG.err; tail -1 gpurun_out/bench_reference_$TAG.json | cut -c1-400
summ() {
  python - "$1" "$2" <<'PY'
import json, sys
W, F = sys.argv[1], sys.argv[2]
try:
    l=json.loads(open(F).read().strip().splitlines()[-1])
    s=l["stats_one_step"]; r=l["roofline"]; nq=s["n_queries"]
    print("%s: value %.3f Mq/s e2e %.3f ms/step %.2f | seed %.2f select %.2f main %.2f | scanned/q %.0f ref %s | parity %s/%s | cpu %s (%s) | roof %s %.3f | clocks %s %s" % (
        W, l["value"]/1e6, l["e2e"]["value"]/1e6, l["ms_per_step"], s["ms_seed_scan"], s["ms_traverse"], s["ms_main_scan"], r["gpu_scanned_points_per_query"],
        r["ref_scanned_points_per_query"], l.get("parity_mismatches"), l.get("parity_checked"), l.get("cpu_baseline",{}).get("value"), l.get("cpu_baseline",{}).get("kind"),
        r["bound"], r["frac"], l["clocks"]["sm_mhz"], l["clocks"]["reasons"]))
except Exception as e:
    print(W, "no line:", e)
PY
}
summ default gpurun_out/bench_default_$TAG.json
for W in posture_like_78095x15_self_K10 gauss_10Mx32_q1M_K10 gauss_10Mx32_q1M_K1 gauss_10Mx32_q1M_K100 siftlike_1Mx128_q100k_K10 gauss_2Mx32_q200k_K10; do
  timeout 900 python bench.py --workload $W --steps 3 --warmup 3 > gpurun_out/bench_${W}_$TAG.json 2> gpurun_out/bench_${W}_$TAG.err
  summ $W gpurun_out/bench_${W}_$TAG.json; tail -2 gpurun_out/bench_${W}_$TAG.err
done
