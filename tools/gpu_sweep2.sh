#!/bin/bash
mkdir -p gpurun_out
echo "== pytest radix"; timeout 900 python -m pytest tests/test_radix_gpu.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -4
: > gpurun_out/variants2.txt
run() { timeout 300 python bench.py --variant $1 --dist $2 --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | tail -1 | python -c "
import sys, json
try:
    j = json.loads(sys.stdin.read()); print('$2', j['config']['radix_variant'], 'ms/sort', round(j['ms_per_step'],3), 'pass_ms', [round(x,3) for x in j['roofline']['kernels']['pass_ms']], 'hist_ms', round(j['roofline']['kernels']['histogram_ms'],3), 'frac', round(j['roofline']['frac'],3))
except Exception as e: print('failed', '$1', '$2', e)
" | tee -a gpurun_out/variants2.txt; }
for v in 0 23 24 25 26; do run $v uniform; done
for d in skewed90 and3 ascending descending all_equal zipf; do run 0 $d; done
