#!/bin/bash
mkdir -p gpurun_out
timeout 240 python tools/variant_check.py 1 > gpurun_out/r02_variant_tma.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_variant_tma.txt
tail -8 gpurun_out/r02_variant_tma.txt
timeout 240 python tools/phase_timing_tma.py > gpurun_out/r02_phase_tma.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_phase_tma.txt
cat gpurun_out/r02_phase_tma.txt
