#!/bin/bash
# round 2, call 1: hardware probes (TMA bulk-store write-out, tensor-memory parking) + baseline bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_probe_gpu.txt 2>&1
timeout 300 ./build/probe_r02 tmem > gpurun_out/r02_probe_tmem.txt 2>&1; echo "tmem rc=$?" >> gpurun_out/r02_probe_tmem.txt
timeout 600 ./build/probe_r02 writeout > gpurun_out/r02_probe_writeout.txt 2>&1; echo "writeout rc=$?" >> gpurun_out/r02_probe_writeout.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_baseline.json 2> gpurun_out/r02_bench_baseline.err
cat gpurun_out/r02_probe_tmem.txt gpurun_out/r02_probe_writeout.txt
tail -c 1500 gpurun_out/r02_bench_baseline.json
