#!/bin/bash
# parity + per-pass timing of one compiled shape (default: the shipped pass kernel, any size), then the phase stamps of
# its timing twin (needs `make experiments`: libb200sort_exp.so).  usage: gpurun -- 'bash tools/gpu_r2_tma3.sh [V] [TWIN]'
V=${1:-3}
TWIN=${2:-TIMING_tma3a}
mkdir -p gpurun_out
timeout 240 python tools/variant_check.py $V > gpurun_out/r02_variant_$V.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_variant_$V.txt
head -1 gpurun_out/r02_variant_$V.txt; grep -v " ok$" gpurun_out/r02_variant_$V.txt | tail -6; tail -2 gpurun_out/r02_variant_$V.txt | head -1
B200SORT_LIB=libb200sort_exp.so timeout 240 python tools/phase_timing_tma.py $TWIN > gpurun_out/r02_phase_$TWIN.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_phase_$TWIN.txt
cat gpurun_out/r02_phase_$TWIN.txt
