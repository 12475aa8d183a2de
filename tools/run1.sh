#!/bin/bash
# first GPU pass of round 2: sanitizer on a small case, the GPU test suite, the full-size tests, the default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -8
nproc; free -g | head -2
echo "== quick functional check"; timeout 300 python tools/sanitize_small.py 2>&1 | tail -12
echo "== sanitizer"; timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/sanitize_r2.txt 2>&1; tail -5 gpurun_out/sanitize_r2.txt
echo "== pytest gpu (without the full-size file)"
timeout 1500 python -m pytest tests -m gpu -q --timeout 400 --timeout-method thread --ignore=tests/test_gpu_fullsize.py -x > gpurun_out/pytest_gpu_r2a.txt 2>&1; tail -25 gpurun_out/pytest_gpu_r2a.txt
echo "== bench default"
timeout 900 python bench.py > gpurun_out/bench_default_r2a.json 2> gpurun_out/bench_default_r2a.err; tail -1 gpurun_out/bench_default_r2a.json | cut -c1-1500; tail -3 gpurun_out/bench_default_r2a.err
echo "== full-size tests"
timeout 2400 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s --timeout 900 --timeout-method thread > gpurun_out/pytest_gpu_fullsize_r2a.txt 2>&1; tail -30 gpurun_out/pytest_gpu_fullsize_r2a.txt
