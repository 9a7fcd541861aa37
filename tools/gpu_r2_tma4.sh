#!/bin/bash
mkdir -p gpurun_out
timeout 240 python tools/phase_timing_tma.py > gpurun_out/r02_phase_tma_nowait.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_phase_tma_nowait.txt
cat gpurun_out/r02_phase_tma_nowait.txt
