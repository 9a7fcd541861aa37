#!/bin/bash
mkdir -p gpurun_out
echo "== atomic order probe"; timeout 120 ./build/atomic_order_probe | tee gpurun_out/atomic_order_probe.txt
echo "== pytest radix with kRankAdd (variant 12)"
timeout 600 python - <<'PY'
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, oracle
from b200sort import datagen
from b200sort._lib import lib, ALGO_RADIX
from helpers import gpu_sort
L = lib()
bad = 0
for v in (12, 13, 14):
    L.b200sort_radix_set_variant(v)
    for dist in ("uniform","and3","skewed90","lab_rand100","edge_mix","mask_00ff00ff","zipf16","all_equal","ascending","descending"):
        for n in (100000, (1<<22)+77):
            if dist.startswith("lab") and n > 200000: continue
            keys = datagen.make(dist, n, 3)
            ok = gpu_sort(keys, ALGO_RADIX).tobytes() == oracle.radix_sort(keys).tobytes()
            bad += (not ok)
            if not ok: print("MISMATCH", v, dist, n)
print("kRankAdd mismatches:", bad)
PY
for v in 12 13 14 3 4; do timeout 300 python bench.py --variant $v --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read()); print(j['config']['radix_variant'], 'ms/sort', round(j['ms_per_step'],3), 'pass_ms', [round(x,3) for x in j['roofline']['kernels']['pass_ms']], 'hist_ms', round(j['roofline']['kernels']['histogram_ms'],3), 'frac', round(j['roofline']['frac'],3))
"; done
CMD="python bench.py --variant 12 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 300 $CMD > gpurun_out/plain_12.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:radix_onesweep -s 8 -c 1 -o gpurun_out/onesweep_v12 $CMD > gpurun_out/ncu_full_12.log 2>&1
echo "ncu exit $?"
