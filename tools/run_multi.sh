#!/bin/bash
# multi-GPU pass: gpurun --gpus N -- bash tools/run_multi.sh N TAG
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}; TAG=${2:-r2m}
nvidia-smi --query-gpu=name --format=csv,noheader | head -8; nvidia-smi topo -m 2>/dev/null | head -12
echo "== pytest: multi handle, drop-in multi, cfg5-shaped sharded"
timeout 900 python -m pytest "tests/test_gpu_multi.py::test_multi_handle_matches_oracle" tests/test_host_cpp.py::test_dropin_subclass_multi_gpu tests/test_gpu_fullsize.py::test_cfg5_shaped_uniform_24Mx16 -m gpu -q --timeout 400 --timeout-method thread -x > gpurun_out/pytest_multi_${TAG}_n$N.txt 2>&1; tail -4 gpurun_out/pytest_multi_${TAG}_n$N.txt
echo "== torchrun bench N=$N (weak scaling line + sharded_dataset + single_process records)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_default_${TAG}_n$N.json 2> gpurun_out/bench_default_${TAG}_n$N.err
python - <<PY
import json
try:
    l=json.loads([x for x in open("gpurun_out/bench_default_${TAG}_n$N.json").read().strip().splitlines() if x.startswith("{")][-1])
    print("value %.3f Mq/s e2e %.3f n_gpus %d parity %s/%s" % (l["value"]/1e6, l["e2e"]["value"]/1e6, l["n_gpus"], l.get("parity_mismatches"), l.get("parity_checked")))
    print("sharded_dataset:", json.dumps(l.get("sharded_dataset")))
    print("single_process:", json.dumps(l.get("single_process")))
    print("extras error:", l.get("multi_gpu_extras_error"))
except Exception as e:
    print("no line:", e)
PY
tail -5 gpurun_out/bench_default_${TAG}_n$N.err
