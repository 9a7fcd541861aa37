#!/bin/bash
mkdir -p gpurun_out
for v in 7 0; do
  timeout 240 python tools/variant_check.py $v > gpurun_out/r02_variant_$v.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_variant_$v.txt
  tail -3 gpurun_out/r02_variant_$v.txt
done
timeout 240 python tools/phase_timing_pp2.py TIMING_pipelined2_prefetch > gpurun_out/r02_phase_pp2_prefetch.txt 2>&1; tail -32 gpurun_out/r02_phase_pp2_prefetch.txt
