#!/bin/bash
# Round-2 evidence on one B200: the GPU test suite, the bench lines, the ncu launch list and full captures of the
# dominant kernels.  usage: gpurun --timeout 3000 -- 'bash tools/gpu_r2_final.sh'
R=r02
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${R}_pytest_gpu.txt
echo "== bench radix"; timeout 900 python bench.py > gpurun_out/${R}_bench_radix.json 2> gpurun_out/${R}_bench_radix.err; echo "rc=$?"
echo "== bench merge"; timeout 600 python bench.py --algo merge --steps 10 --no-cpu-baseline --no-configs > gpurun_out/${R}_bench_merge.json 2>/dev/null; echo "rc=$?"
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>/dev/null; echo "rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-configs"
timeout 300 $CMD > gpurun_out/${R}_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'radix|merge|block_sort|dist_' -c 70 --csv --log-file gpurun_out/${R}_launches.csv $CMD > gpurun_out/${R}_ncu_launches.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:radix_onesweep -s 8 -c 1 -f -o gpurun_out/${R}_onesweep $CMD > gpurun_out/${R}_ncu_onesweep.log 2>&1
echo "onesweep exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:radix_histogram -s 2 -c 1 -f -o gpurun_out/${R}_hist $CMD > gpurun_out/${R}_ncu_hist.log 2>&1
echo "hist exit $?"
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02_bench_radix.json').read().strip().splitlines()[-1])
print('radix: Gkeys/s', round(j['value']/1e9,2), 'ms', round(j['ms_per_step'],4), 'frac', round(j['roofline']['frac'],4), 'e2e', round(j['e2e']['value']/1e9,2), 'cpu', j['cpu_baseline']['value']/1e6 if j['cpu_baseline'] else None)
for r in j['configs'] or []:
    print('  ', r.get('config'), 'ms', round(r.get('ms_per_step', -1), 3), 'Gk/s', round(r.get('keys_per_s', 0)/1e9, 1), 'pass_frac', round(r.get('pass_frac') or 0, 3), r.get('error', ''))
m=json.loads(open('gpurun_out/r02_bench_merge.json').read().strip().splitlines()[-1])
print('merge: Gkeys/s', round(m['value']/1e9,2), 'ms', round(m['ms_per_step'],3), m['roofline']['kernels'])
r=json.loads(open('gpurun_out/r02_bench_reference.json').read().strip().splitlines()[-1])
print('reference arm: Mkeys/s', round(r['value']/1e6,1))
PY
