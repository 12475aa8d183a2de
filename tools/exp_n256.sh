#!/bin/bash
# GPU box only: whole-page MMA chains (SCHT_TC_N256=1, two issuers) against the default on wide and narrow points
cd "$(dirname "$0")/.."
run() {  # env..., workload
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload $1 2>/dev/null | python -c "
import json,sys,os
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stats_one_step']
print(sys.argv[1], 'N256=%s ISS=%s' % (os.environ.get('SCHT_TC_N256','-'), os.environ.get('SCHT_TC_ISSUERS','-')), round(d['value']), round(d['ms_per_step'],2), {k:round(s[k],2) for k in ('ms_seed_scan','ms_main_scan')})
" $1
}
for w in siftlike_1Mx128_q100k_K10 gauss_2Mx32_q200k_K10 uniform_1Mx16_q100k_K10; do
  run $w
  SCHT_TC_ISSUERS=2 run $w
  SCHT_TC_ISSUERS=2 SCHT_TC_N256=1 run $w
done
SCHT_TC_ISSUERS=2 SCHT_TC_N256=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "prefilter or sample_reuse or full_size or random_vs_oracle" 2>&1 | tail -2
