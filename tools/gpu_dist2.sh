#!/bin/bash
# usage gpu_dist2.sh N : scaling bench at N GPUs (p2p + nccl), skewed run, plus host-path check on GPU 0
N=${1:-4}
mkdir -p gpurun_out
echo "== lab tests (host operator incl. staged pageable path)"; timeout 900 python -m pytest tests/test_lab_gpu.py tests/test_merge_gpu.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -4
echo "== checked driver (pageable host arrays)"; (cd gpurun_out && ../build/b200sort_driver --min 16777216 --max 67108864 --dist uniform --check --csv drv.csv | tail -9)
echo "== dist check x$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_2gpu_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -10
for ex in p2p nccl; do
echo "== bench x$N $ex"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --exchange $ex 2>&1 | grep -v "^W\|^\*\*\*" | tail -1 | tee gpurun_out/bench_dist_${N}_$ex.json | python -c "
import sys, json
j = json.loads(sys.stdin.read()); print('N=', j['n_gpus'], '$ex', 'ms/step', round(j['ms_per_step'],3), 'Gkeys/s', round(j['value']/1e9,1), 'e2e Gkeys/s', round(j['e2e']['value']/1e9,2), 'recv', j['config']['recv_counts'])"
done
echo "== bench x$N p2p skewed90"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --dist skewed90 --log2n 26 2>&1 | grep -v "^W\|^\*\*\*" | tail -1 | tee gpurun_out/bench_dist_${N}_skewed.json | python -c "
import sys, json
j = json.loads(sys.stdin.read()); print('skewed90 N=', j['n_gpus'], 'ms/step', round(j['ms_per_step'],3), 'Gkeys/s', round(j['value']/1e9,1), 'recv', j['config']['recv_counts'])"
