#!/bin/bash
# 8-GPU box: weak-scaling line (2^28 keys per GPU) and BASELINE config 5 (n = 2^30 in total: 2^27 per GPU at P = 8)
N=${1:-8}
mkdir -p gpurun_out
for L in 28 27; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$L bench.py --gpus $N --steps 10 --warmup 3 --log2n $L --e2e-steps 2 2>&1 | grep -v "^W\|^\*\*\*" | tail -1 | tee gpurun_out/r01_bench_dist_${N}gpu_p2p_log2n$L.json | python -c "
import sys, json
j = json.loads(sys.stdin.read()); print('N=', j['n_gpus'], 'log2n/GPU', $L, 'ms/step', round(j['ms_per_step'],3), 'Gkeys/s', round(j['value']/1e9,1), {k: round(v,3) for k,v in j['roofline']['phases_max_over_ranks'].items()}, 'nvlink out GB/s', round(j['roofline']['nvlink_gbs_per_gpu_out'],1), 'e2e', round(j['e2e']['value']/1e9,2))"
done
