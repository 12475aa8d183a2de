#!/bin/bash
# GPU box: value-carrying ring entries + re-check (tc_recheck) against ring depth
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -x -q -m gpu --timeout 200 --timeout-method thread 2>&1 | tail -2
for W in gauss_2Mx32_q200k_K10 siftlike_1Mx128_q100k_K10 gauss_10Mx32_q1M_K10 gauss_2Mx32_q200k_K100 gauss_2Mx32_q200k_K1 uniform_1Mx16_q100k_K10; do
  timeout 500 python tools/exp_opts.py $W tc_recheck=0 - tc_ring=256 tc_ring=32 2>&1 | grep -v "^$" | tail -4
done
