#!/bin/bash
# gpurun --gpus 8 -- bash tools/run_cfg5.sh TAG
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2m}
echo "== cfg5: uniform 100M x 16, 10M queries, one process, 8 GPUs behind one handle (both sharding modes)"
timeout 900 python bench.py --single-process --gpus 8 --workload uniform_100Mx16_q10M_K10 --steps 1 --warmup 0 --cpu-sample 48 > gpurun_out/bench_cfg5_single_process_${TAG}_n8.json 2> gpurun_out/bench_cfg5_single_process_${TAG}_n8.err
python - <<PY
import json
try:
    l=json.loads([x for x in open("gpurun_out/bench_cfg5_single_process_${TAG}_n8.json").read().strip().splitlines() if x.startswith("{")][-1])
    print("cfg5 value %.3f Mq/s parity %s/%s clocks %s" % (l["value"]/1e6, l.get("parity_mismatches"), l.get("parity_checked"), l["clocks"]))
    print(json.dumps(l["single_process"])); print(json.dumps(l.get("tree")), json.dumps(l.get("cpu_baseline")))
except Exception as e:
    print("no line:", e)
PY
tail -5 gpurun_out/bench_cfg5_single_process_${TAG}_n8.err
