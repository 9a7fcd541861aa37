#!/bin/bash
# usage: gpu_r2_dist.sh N   (under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_${N}.txt 2>&1
echo "== multi-device tests"; timeout 1200 python -m pytest tests/test_multi_device_gpu.py -x -q 2>&1 | tail -6
if [ "$N" -ge 4 ]; then
  echo "== bench weak x$N, digit histograms by the local sort's own kernel (B200SORT_DIST_HIST_AT_SOURCE=0)"
  B200SORT_DIST_HIST_AT_SOURCE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --steps 20 --warmup 3 --no-configs 2>/dev/null | tail -1 > gpurun_out/r02_bench_dist_${N}_hist_local.json
  python -c "
import json; j=json.loads(open('gpurun_out/r02_bench_dist_${N}_hist_local.json').read()); print('  ms', round(j['ms_per_step'],3), 'Gkeys/s', round(j['value']/1e9,1), {k: round(v,3) for k,v in j['roofline']['phases_max_over_ranks'].items()})"
fi
echo "== bench weak x$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 2> gpurun_out/r02_bench_dist_${N}.err | tail -1 > gpurun_out/r02_bench_dist_${N}.json
python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/r02_bench_dist_${N}.json').read())
    print('weak x${N}: ms', round(j['ms_per_step'],3), 'Gkeys/s', round(j['value']/1e9,1), 'phases', {k: round(v,3) for k,v in j['roofline']['phases_max_over_ranks'].items()}, 'nvlink GB/s out', round(j['roofline']['nvlink_gbs_per_gpu_out'],0), 'e2e', round(j['e2e']['value']/1e9,2))
    for r in j.get('configs') or []:
        print('  ', r.get('config'), 'ms', round(r.get('ms_per_step',-1),3), 'Gk/s', round(r.get('keys_per_s',0)/1e9,1), 'max/mean', r.get('recv_max_over_mean'), r.get('error',''), {k: round(v,3) for k,v in (r.get('phases_max_over_ranks') or {}).items()})
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r02_bench_dist_${N}.err').read()[-3000:])
PY
