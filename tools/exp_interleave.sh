#!/bin/bash
# GPU box only: secondary workloads + default-workload A/B after a k_scan_tc change, then the parity tests of test_gpu_parity.py
cd "$(dirname "$0")/.."
for w in gauss_2Mx32_q200k_K100 gauss_2Mx32_q200k_K10 siftlike_1Mx128_q100k_K10 posture_like_78095x15_self_K10; do
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload $w 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stats_one_step']
print(sys.argv[1], round(d['value']), round(d['ms_per_step'],2), {k:round(s[k],2) for k in ('ms_seed_scan','ms_main_scan')}, s['exact_reevaluations'])
" $w
done
python tools/ab_cfg2.py tools/bin/libscht_old.so semi-convex_hull_tree_b200/libscht_b200.so 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -1
