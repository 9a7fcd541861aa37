#!/bin/bash
mkdir -p gpurun_out
timeout 600 ./build/probe_tma_scatter > gpurun_out/r02_probe_tma_scatter.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_probe_tma_scatter.txt
cat gpurun_out/r02_probe_tma_scatter.txt
