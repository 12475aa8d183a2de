# usage: bash tools/seed_sweep.sh  -- effect of the seed neighbourhood size (SCHT_SEED_POINTS) per workload and scan path
for tc in ${TC_LIST:-2 0}; do
for w in ${WORKLOAD_LIST:-posture_like_78095x15_self_K10 siftlike_1Mx128_q100k_K10 gauss_2Mx32_q200k_K10 gauss_2Mx32_q200k_K100}; do
 for sp in ${SEED_LIST:-4096 1024 512}; do
  SCHT_TC=$tc SCHT_SEED_POINTS=$sp timeout ${BENCH_TIMEOUT:-280} python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
s=d['stats_one_step']
print('tc=$tc seed=$sp', d['config']['workload'], round(d['value']), 'ms', round(d['ms_per_step'],2), 'seed', round(s['ms_seed_scan'],2), 'trav', round(s['ms_traverse'],2), 'main', round(s['ms_main_scan'],2), 'tc', s['prefilter_batches'], 'exact', s['exact_reevaluations'])
"
 done
done
done
