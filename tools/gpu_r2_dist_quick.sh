#!/bin/bash
# usage: gpu_r2_dist_quick.sh N   (under gpurun --gpus N): the distributed parity tests and one weak-scaling bench line
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist_gpu.py -x -q 2>&1 | tail -1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 3 --no-configs 2>/dev/null | tail -1 > gpurun_out/r02_bench_dist_${N}_quick.json
python -c "
import json; j=json.loads(open('gpurun_out/r02_bench_dist_${N}_quick.json').read()); print('weak x$N: ms', round(j['ms_per_step'],3), 'Gkeys/s', round(j['value']/1e9,1), {k: round(v,3) for k,v in j['roofline']['phases_max_over_ranks'].items()})"
