#!/bin/bash
# GPU box: vectorised k_descend -- parity subset and device times
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -x -q -m gpu --timeout 200 --timeout-method thread 2>&1 | tail -2
for W in uniform_1Mx16_q100k_K10 gauss_2Mx32_q200k_K10 siftlike_1Mx128_q100k_K10 gauss_10Mx32_q1M_K10; do
  timeout 500 python tools/exp_opts.py $W - 2>&1 | grep -v "^$" | tail -1
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_descend -c 6 --csv python tools/exp_opts.py gauss_2Mx32_q200k_K10 - 2>/dev/null | grep k_descend | tail -2 | cut -d, -f5,15-
