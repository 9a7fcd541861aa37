#!/bin/bash
mkdir -p gpurun_out
for v in 0 7 8 5; do
  timeout 240 python tools/variant_check.py $v > gpurun_out/r02_variant_$v.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_variant_$v.txt
  head -1 gpurun_out/r02_variant_$v.txt; tail -3 gpurun_out/r02_variant_$v.txt
done
