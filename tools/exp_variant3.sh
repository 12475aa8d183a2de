#!/bin/bash
# GPU box only: k_scan with 4 resident CTAs per SM (96 registers, 2-stage ring) against the default 3
cd "$(dirname "$0")/.."
for w in uniform_1Mx16_q100k_K10 siftlike_1Mx128_q100k_K10 gauss_2Mx32_q200k_K100; do
 for v in 2 3; do
  SCHT_SCAN_VARIANT=$v python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload $w 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stats_one_step']
print(sys.argv[1], 'variant', sys.argv[2], round(d['value']), round(d['ms_per_step'],2), {k:round(s[k],2) for k in ('ms_seed_scan','ms_main_scan')})
" $w $v
 done
done
SCHT_SCAN_VARIANT=3 python tools/quick_cfg2.py --check 2>&1 | tail -1
SCHT_SCAN_VARIANT=3 timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "random_vs_oracle or edge or duplicate or golden" 2>&1 | tail -1
