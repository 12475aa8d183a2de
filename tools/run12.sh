#!/bin/bash
# GPU box: survivor ring depth (tc_ring) of k_scan_tc
cd "$(dirname "$0")/.."
for W in gauss_2Mx32_q200k_K10 siftlike_1Mx128_q100k_K10 gauss_10Mx32_q1M_K10 uniform_1Mx16_q100k_K10 gauss_2Mx32_q200k_K100; do
  timeout 500 python tools/exp_opts.py $W - tc_ring=32 tc_ring=64 tc_ring=128 2>&1 | grep -v "^$" | tail -4
done
