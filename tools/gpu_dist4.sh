#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
echo "== dist kernel tests (both shapes)"; timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -2
B200SORT_DIST_SHAPE=1 timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -2
for shape in 0 1; do
B200SORT_DIST_SHAPE=$shape timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 2>&1 | grep -v "^W\|^\*\*\*" | tail -1 | tee gpurun_out/bench_dist_${N}_shape$shape.json | python -c "
import sys, json
j = json.loads(sys.stdin.read()); print('shape $shape N=', j['n_gpus'], 'ms/step', round(j['ms_per_step'],3), 'Gkeys/s', round(j['value']/1e9,1), {k: round(v,3) for k,v in j['roofline']['phases_max_over_ranks'].items()}, 'nvlink out GB/s', round(j['roofline']['nvlink_gbs_per_gpu_out'],1))"
done
