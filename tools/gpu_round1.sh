#!/bin/bash
# First GPU trip: smoke, parity tests, bench (radix + merge), tile-shape sweep, ncu launch list + one full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== bench radix"; timeout 600 python bench.py --steps 50 --warmup 5 2>&1 | tail -3 | tee gpurun_out/bench_radix.json
echo "== bench merge"; timeout 600 python bench.py --algo merge --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | tail -3 | tee gpurun_out/bench_merge.json
echo "== variants"
for v in 1 2 3 4 5 6 7; do timeout 300 python bench.py --variant $v --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | tail -1 | python -c "
import sys, json
try:
    j = json.loads(sys.stdin.read()); print(j['config']['radix_variant'], 'ms/sort', round(j['ms_per_step'],3), 'pass_ms', [round(x,3) for x in j['roofline']['kernels']['pass_ms']], 'hist_ms', round(j['roofline']['kernels']['histogram_ms'],3), 'frac', round(j['roofline']['frac'],3))
except Exception as e: print('variant failed', e)
" | tee -a gpurun_out/variants.txt; done
echo "== ncu launches"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'radix|merge|block_sort' -c 80 --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:radix_onesweep -s 8 -c 2 -o gpurun_out/onesweep_r01 $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:radix_histogram -s 2 -c 1 -o gpurun_out/hist_r01 $CMD > gpurun_out/ncu_full_hist.log 2>&1
echo "ncu hist exit $?"
