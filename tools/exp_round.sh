#!/bin/bash
# GPU box only: A/B on the default workload against tools/bin/libscht_old.so, the secondary workloads, and the prefilter parity tests
cd "$(dirname "$0")/.."
python tools/ab_cfg2.py tools/bin/libscht_old.so semi-convex_hull_tree_b200/libscht_b200.so 2>&1
for w in ${WORKLOADS:-siftlike_1Mx128_q100k_K10 gauss_2Mx32_q200k_K10 gauss_2Mx32_q200k_K100}; do
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload $w 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stats_one_step']
print(sys.argv[1], round(d['value']), round(d['ms_per_step'],2), {k:round(s[k],2) for k in ('ms_seed_scan','ms_main_scan')}, s['exact_reevaluations'], s['list_insertions'])
" $w
done
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
