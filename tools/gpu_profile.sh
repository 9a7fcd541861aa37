#!/bin/bash
# Round profile: launch list of one bench run + ncu --set full of the dominant kernels. usage: gpu_profile.sh rNN
R=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 300 $CMD > gpurun_out/${R}_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'radix|merge|block_sort|dist_' -c 70 --csv --log-file gpurun_out/${R}_launches.csv $CMD > gpurun_out/${R}_ncu_launches.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:radix_onesweep -s 8 -c 1 -o gpurun_out/${R}_onesweep $CMD > gpurun_out/${R}_ncu_onesweep.log 2>&1
echo "onesweep exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:radix_histogram -s 2 -c 1 -o gpurun_out/${R}_hist $CMD > gpurun_out/${R}_ncu_hist.log 2>&1
echo "hist exit $?"
CMDM="python bench.py --algo merge --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 300 $CMDM > gpurun_out/${R}_plain_merge.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'merge_pass' -s 12 -c 1 -o gpurun_out/${R}_merge $CMDM > gpurun_out/${R}_ncu_merge.log 2>&1
echo "merge pass exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'block_sort' -s 2 -c 1 -o gpurun_out/${R}_blocksort $CMDM > gpurun_out/${R}_ncu_blocksort.log 2>&1
echo "block sort exit $?"
echo "== full bench (radix, default K/W)"; timeout 600 python bench.py 2>&1 | tail -1 | tee gpurun_out/${R}_bench_radix.json
echo "== full bench merge"; timeout 600 python bench.py --algo merge --steps 10 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${R}_bench_merge.json
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | tee gpurun_out/${R}_bench_reference.json
echo "== dists"; : > gpurun_out/${R}_dists.txt; for d in and3 mask_0000ffff skewed90 ascending descending; do timeout 300 python bench.py --dist $d --steps 20 --no-cpu-baseline --e2e-steps 1 2>&1 | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read()); print('$d', 'ms/sort', round(j['ms_per_step'],3), 'Gkeys/s', round(j['value']/1e9,1), 'pass_ms', [round(x,3) for x in j['roofline']['kernels']['pass_ms']], 'hist_ms', round(j['roofline']['kernels']['histogram_ms'],3))
" | tee -a gpurun_out/${R}_dists.txt; done
echo "== dists, merge"; for d in and3 skewed90 ascending descending; do timeout 300 python bench.py --algo merge --dist $d --steps 5 --no-cpu-baseline --e2e-steps 1 2>&1 | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read()); print('merge $d', 'ms/sort', round(j['ms_per_step'],3), 'Gkeys/s', round(j['value']/1e9,1), j['roofline']['kernels'])
" | tee -a gpurun_out/${R}_dists.txt; done
echo "== drivers"; (cd gpurun_out && ../build/sort && cat output.txt | head -20 > ${R}_driver_main_output.txt; ../build/performaceTest > ${R}_driver_perftest.txt; ../build/b200sort_driver --min 256 --max 67108864 --dist uniform --check --csv ${R}_driver_checked.csv > ${R}_driver_checked.txt; tail -25 ${R}_driver_checked.txt)
