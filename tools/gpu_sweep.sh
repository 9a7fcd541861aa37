#!/bin/bash
# Radix tile-shape / rank-mode sweep + ncu captures.  usage: gpu_sweep.sh "<variants>" "<ncu variants>"
mkdir -p gpurun_out
VARS=${1:-"0 1 2 3 4 5 6 7 8 9 10 11"}
NCUV=${2:-"0 1"}
echo "== pytest radix"; timeout 900 python -m pytest tests/test_radix_gpu.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -5
echo "== variants"
: > gpurun_out/variants.txt
for v in $VARS; do timeout 300 python bench.py --variant $v --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | tail -1 | python -c "
import sys, json
try:
    j = json.loads(sys.stdin.read()); print(j['config']['radix_variant'], 'ms/sort', round(j['ms_per_step'],3), 'pass_ms', [round(x,3) for x in j['roofline']['kernels']['pass_ms']], 'hist_ms', round(j['roofline']['kernels']['histogram_ms'],3), 'frac', round(j['roofline']['frac'],3))
except Exception as e: print('variant failed', e)
" | tee -a gpurun_out/variants.txt; done
for v in $NCUV; do
CMD="python bench.py --variant $v --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 300 $CMD > gpurun_out/plain_$v.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:radix_onesweep -s 8 -c 1 -o gpurun_out/onesweep_v$v $CMD > gpurun_out/ncu_full_$v.log 2>&1
echo "ncu variant $v exit $?"
done
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:radix_histogram -s 2 -c 1 -o gpurun_out/hist_r02 $CMD > gpurun_out/ncu_full_hist.log 2>&1
echo "ncu hist exit $?"
