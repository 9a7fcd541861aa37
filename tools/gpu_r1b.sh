#!/bin/bash
# Round-1 second session: A/B of the split/packed onesweep shapes and of the new merge kernels.
# usage: gpu_r1b.sh "<radix variants>" "<merge variants>" [ncu kernel regex] [ncu launches to skip]
mkdir -p gpurun_out
echo "== pytest (all onesweep shapes, merge)"
timeout 900 python -m pytest tests/test_radix_gpu.py -m gpu -q -x --timeout 300 -p no:cacheprovider -k "all_tile_shapes or single_pass" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_merge_gpu.py -m gpu -q -x --timeout 600 -p no:cacheprovider 2>&1 | tail -3
: > gpurun_out/variants_r1b.txt
for v in $1; do timeout 300 python bench.py --variant $v --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | tail -1 | python -c "
import sys, json
try:
    j = json.loads(sys.stdin.read()); print($v, j['config']['radix_variant'], 'ms/sort', round(j['ms_per_step'],3), 'pass_ms', [round(x,3) for x in j['roofline']['kernels']['pass_ms']], 'hist_ms', round(j['roofline']['kernels']['histogram_ms'],3), 'frac', round(j['roofline']['frac'],3))
except Exception as e: print('variant failed', $v, e)
" | tee -a gpurun_out/variants_r1b.txt; done
for m in $2; do timeout 300 python bench.py --algo merge --merge-variant $m --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | tail -1 | python -c "
import sys, json
try:
    j = json.loads(sys.stdin.read()); print('merge', $m, 'ms/sort', round(j['ms_per_step'],3), 'Gkeys/s', round(j['value']/1e9,2), j['roofline']['kernels'], 'pass frac', round(j['roofline']['frac'],3))
except Exception as e: print('merge variant failed', $m, e)
" | tee -a gpurun_out/variants_r1b.txt; done
if [ -n "$3" ]; then
  CMDM="python bench.py --algo merge --log2n 26 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$3" -s ${4:-24} -c 2 -o gpurun_out/r01b_merge $CMDM > gpurun_out/r01b_ncu_merge.log 2>&1
  echo "ncu merge exit $?"
fi
