#!/bin/bash
# usage: gpu_sweep3.sh "<variants>" "<timing name filter or empty>"
mkdir -p gpurun_out
echo "== pytest radix"; timeout 900 python -m pytest tests/test_radix_gpu.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -3
: > gpurun_out/variants.txt
for v in $1; do timeout 300 python bench.py --variant $v --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | tail -1 | python -c "
import sys, json
try:
    j = json.loads(sys.stdin.read()); print($v, j['config']['radix_variant'], 'ms/sort', round(j['ms_per_step'],3), 'pass_ms', [round(x,3) for x in j['roofline']['kernels']['pass_ms']], 'hist_ms', round(j['roofline']['kernels']['histogram_ms'],3), 'frac', round(j['roofline']['frac'],3))
except Exception as e: print('variant failed', $v, e)
" | tee -a gpurun_out/variants.txt; done
if [ -n "$2" ]; then timeout 300 python tools/phase_timing.py "$2" 2>&1 | tee gpurun_out/phase_timing_$2.txt; fi
