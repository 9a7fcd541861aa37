#!/bin/bash
# parity + per-pass timing of one compiled shape (default: the 16384-key TMA kernel), then its phase stamps
V=${1:-3}
mkdir -p gpurun_out
timeout 240 python tools/variant_check.py $V > gpurun_out/r02_variant_$V.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_variant_$V.txt
head -1 gpurun_out/r02_variant_$V.txt; grep -v " ok$" gpurun_out/r02_variant_$V.txt | tail -6
timeout 240 python tools/phase_timing_tma.py ${2:-TIMING_tma2} > gpurun_out/r02_phase_tma2.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_phase_tma2.txt
cat gpurun_out/r02_phase_tma2.txt
