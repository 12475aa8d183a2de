#!/bin/bash
# GPU box: pause of an epilogue warp whose survivor ring is full
cd "$(dirname "$0")/.."
for W in gauss_2Mx32_q200k_K10 siftlike_1Mx128_q100k_K10 gauss_10Mx32_q1M_K10 gauss_2Mx32_q200k_K100 uniform_1Mx16_q100k_K10; do
  timeout 500 python tools/exp_opts.py $W - ring_wait_ns=200 ring_wait_ns=1000 ring_wait_ns=4000 2>&1 | grep -v "^$" | tail -4
done
