#!/bin/bash
mkdir -p gpurun_out
echo "== pytest merge"; timeout 900 python -m pytest tests/test_merge_gpu.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -3
for a in merge lab; do timeout 600 python bench.py --algo $a --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read()); print('$a', 'ms/sort', round(j['ms_per_step'],3), 'Gkeys/s', round(j['value']/1e9,2), j['roofline']['kernels'], 'pass frac', round(j['roofline']['frac'],3))"; done
