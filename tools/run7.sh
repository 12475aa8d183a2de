#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02a}
echo "== single-process path on one GPU (smoke workload, both modes)"
timeout 300 python bench.py --single-process --gpus 1 --workload smoke_50kx16_q4k_K10 --steps 2 --warmup 0 --cpu-sample 32 2>&1 | tail -3 | cut -c1-1200
echo "== e2e with a split single batch"
for SP in 0 2 3; do SCHT_PIPELINE_SPLIT=$SP timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pipeline_split $SP: value %.3f e2e %.3f Mq/s' % (l['value']/1e6, l['e2e']['value']/1e6))"; done
echo "== ncu: cfg2"
bash tools/profile.sh ${TAG}_cfg2 uniform_1Mx16_q100k_K10 > gpurun_out/profile_${TAG}_cfg2.log 2>&1; tail -3 gpurun_out/profile_${TAG}_cfg2.log
echo "== ncu: cfg3 (2M x 32 stand-in of the clustered family: same kernels, shorter run)"
bash tools/profile.sh ${TAG}_gauss2M gauss_2Mx32_q200k_K10 k_select_pages > gpurun_out/profile_${TAG}_gauss2M.log 2>&1; tail -3 gpurun_out/profile_${TAG}_gauss2M.log
rm -f gpurun_out/ncu_*_${TAG}_*.log
ls -la gpurun_out/*.ncu-rep | tail
