#!/bin/bash
# GPU box only: what the driver runs at round end -- GPU tests, smoke(), the bench line (with the CPU baseline) and the reference arm
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.txt 2>&1; tail -2 gpurun_out/pytest_gpu_final.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -1 gpurun_out/bench_default.json | cut -c1-330
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -1 gpurun_out/bench_reference.json | cut -c1-400
