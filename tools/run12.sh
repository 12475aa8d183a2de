#!/bin/bash
# GPU box: lazy seed scan as a template instantiation -- new test, parity subset, device times
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -x -q -m gpu --timeout 200 --timeout-method thread 2>&1 | tail -2
for W in uniform_1Mx16_q100k_K10 gauss_2Mx32_q200k_K10 siftlike_1Mx128_q100k_K10 gauss_10Mx32_q1M_K10; do
  timeout 500 python tools/exp_opts.py $W lazy_seed=0 lazy_seed=1 2>&1 | grep -v "^$" | tail -2
done
