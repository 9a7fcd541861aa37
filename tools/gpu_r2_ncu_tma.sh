#!/bin/bash
# one `ncu --set full` capture of the pass kernel of shape V (default 3: the shipped kernel at any size), after the same
# command has exited 0 without ncu.  usage: gpurun -- 'bash tools/gpu_r2_ncu_tma.sh [V]'
mkdir -p gpurun_out
CMD="python tools/run_variant_sort.py ${1:-3} 28 2"
timeout 300 $CMD > gpurun_out/r02_plain_tma.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:radix_onesweep_tma -s 5 -c 1 -f -o gpurun_out/r02_onesweep_tma3 $CMD > gpurun_out/r02_ncu_tma.log 2>&1
echo "ncu exit $?"; tail -5 gpurun_out/r02_ncu_tma.log
