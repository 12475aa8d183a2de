#!/bin/bash
# GPU box only: 2-stage against 4-stage slice ring of the seed scan on wide points
cd "$(dirname "$0")/.."
for st in 4 0; do
  SCHT_SCAN_STAGES=$st python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload siftlike_1Mx128_q100k_K10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stats_one_step']
print('siftlike stages', sys.argv[1], round(d['value']), round(d['ms_per_step'],2), {k:round(s[k],2) for k in ('ms_seed_scan','ms_main_scan')})
" $st
done
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
