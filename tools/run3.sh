#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2c}
echo "== wide points (each case in its own process, 60 s limit)"
for D in 144 160 192 256 512; do for TC in 0 1; do timeout 60 python tools/exp_wide.py $D $TC 2>&1 | tail -1 || echo "dim $D tc $TC: TIMEOUT/FAIL rc=$?"; done; done
echo "== quick functional check"; timeout 300 python tools/sanitize_small.py 2>&1 | tail -12
echo "== bench default"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_default_$TAG.json 2> gpurun_out/bench_default_$TAG.err; tail -3 gpurun_out/bench_default_$TAG.err
for W in uniform_1Mx16_q100k_K10 gauss_2Mx32_q200k_K10 gauss_2Mx32_q200k_K100 siftlike_1Mx128_q100k_K10 gauss_10Mx32_q1M_K10; do
  [ $W != uniform_1Mx16_q100k_K10 ] && timeout 600 python bench.py --workload $W --steps 3 --warmup 3 --cpu-sample 64 > gpurun_out/bench_${W}_$TAG.json 2> gpurun_out/bench_${W}_$TAG.err
  F=gpurun_out/bench_${W}_$TAG.json; [ $W = uniform_1Mx16_q100k_K10 ] && F=gpurun_out/bench_default_$TAG.json
  python - <<PY
import json
try:
    l=json.loads(open("$F").read().strip().splitlines()[-1])
    s=l["stats_one_step"]; r=l["roofline"]; t=l["tree"]
    print("$W: value %.3f Mq/s e2e %.3f ms/step %.2f | seed %.2f select %.2f main %.2f | scanned/query gpu %.0f ref %s | tc %d | parity %s/%s | create %.2fs groups %d supers %d pages %d | clocks %s %s" % (
        l["value"]/1e6, l["e2e"]["value"]/1e6, l["ms_per_step"], s["ms_seed_scan"], s["ms_traverse"], s["ms_main_scan"], r["gpu_scanned_points_per_query"],
        r["ref_scanned_points_per_query"], s["prefilter_batches"], l.get("parity_mismatches"), l.get("parity_checked"), t["device_create_s"], t["scan_groups"], t["super_groups"], t["scan_pages"],
        l["clocks"]["sm_mhz"], l["clocks"]["reasons"]))
except Exception as e:
    print("$W: no line:", e)
PY
  tail -2 gpurun_out/bench_${W}_$TAG.err 2>/dev/null
done
echo "== pytest gpu (multi + parity files)"
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py tests/test_sharding.py -m gpu -q --timeout 120 --timeout-method thread -x --deselect tests/test_gpu_multi.py::test_prefilter_margin_wide_points > gpurun_out/pytest_gpu_$TAG.txt 2>&1; tail -8 gpurun_out/pytest_gpu_$TAG.txt
