#!/bin/bash
# Round-1: 3-CTA/SM onesweep shapes, sort-by-key tile size A/B, BASELINE config 2 (n = 2^24) table.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_radix_gpu.py -m gpu -q -x --timeout 300 -p no:cacheprovider -k "all_tile_shapes" 2>&1 | tail -2
B200SORT_PAIRS_IPT=12 timeout 600 python -m pytest tests/test_radix_gpu.py -m gpu -q -x --timeout 300 -p no:cacheprovider -k "sort_by_key" 2>&1 | tail -2
for v in 53 54; do timeout 300 python bench.py --variant $v --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | tail -1 | python -c "
import sys, json
try:
    j = json.loads(sys.stdin.read()); print($v, j['config']['radix_variant'], 'ms/sort', round(j['ms_per_step'],3), 'pass_ms', [round(x,3) for x in j['roofline']['kernels']['pass_ms']])
except Exception as e: print('variant failed', $v, e)
"; done
python tools/pairs_timing.py 28 20; B200SORT_PAIRS_IPT=12 python tools/pairs_timing.py 28 20
echo "== BASELINE config 2: n = 2^24, radix, uniform and low-entropy keys (seed 1)" | tee gpurun_out/r01_config2_n24.txt
for d in uniform and2 and3 and4 mask_0000ffff mask_00ff00ff; do timeout 300 python bench.py --log2n 24 --dist $d --steps 50 --no-cpu-baseline --e2e-steps 1 2>&1 | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read()); print('$d', 'ms/sort', round(j['ms_per_step'],4), 'Gkeys/s', round(j['value']/1e9,1), 'pass_ms', [round(x,3) for x in j['roofline']['kernels']['pass_ms']], 'hist_ms', round(j['roofline']['kernels']['histogram_ms'],3), 'e2e_ms', round(j['e2e']['ms_per_step'],2))
" | tee -a gpurun_out/r01_config2_n24.txt; done
