#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_bench_radix.json 2> gpurun_out/r02_bench_radix.err; echo "rc=$?"
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02_bench_radix.json').read().strip().splitlines()[-1])
print('value Gkeys/s', j['value']/1e9, 'ms', j['ms_per_step'], 'frac', j['roofline']['frac'], 'e2e', j['e2e']['value']/1e9, 'traffic', j['roofline']['traffic'])
for r in j['configs'] or []:
    print(r.get('config'), 'ms', round(r.get('ms_per_step', -1), 3), 'Gk/s', round(r.get('keys_per_s', 0)/1e9, 1), 'pass_frac', round(r.get('pass_frac') or 0, 3), r.get('error', ''))
PY
tail -3 gpurun_out/r02_bench_radix.err
