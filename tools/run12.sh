#!/bin/bash
# GPU box: first-pass rule with the re-checking rescorers
cd "$(dirname "$0")/.."
for W in siftlike_1Mx128_q100k_K10 gauss_2Mx32_q200k_K10 gauss_10Mx32_q1M_K10 gauss_2Mx32_q200k_K100; do
  timeout 500 python tools/exp_opts.py $W - first_pass=0 first_pass=1 first_pass=2 2>&1 | grep -v "^$" | tail -4
done
