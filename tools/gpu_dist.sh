#!/bin/bash
# Multi-GPU trip: usage gpu_dist.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
echo "== single-GPU dist kernel tests"; timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -4
echo "== dist check x$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_2gpu_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -14
for ex in p2p nccl; do
echo "== bench x$N $ex"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --exchange $ex 2>&1 | grep -v "^W\|^\*\*\*" | tail -2 | tee gpurun_out/bench_dist_${N}_$ex.json
done
echo "== bench x1 (scaling denominator)"; timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 2 2>&1 | tail -1 | tee gpurun_out/bench_1.json
