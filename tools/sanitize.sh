#!/bin/bash
# compute-sanitizer over every kernel family at smoke sizes.  ONE tool per gpurun call (B200_PROFILING.md):
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh memcheck'      (also: racecheck | synccheck | initcheck)
# Logs: gpurun_out/r02_sanitizer_<tool>.txt (copied to profiles/).
TOOL=${1:-memcheck}
mkdir -p gpurun_out
OUT=gpurun_out/r02_sanitizer_${TOOL}.txt
ARGS=""
[ "$TOOL" = "memcheck" ] && ARGS="big"
[ "$TOOL" = "initcheck" ] && ARGS="big"
echo "# python tools/sanitize_target.py $ARGS (plain run first)" > $OUT
timeout 600 python tools/sanitize_target.py $ARGS >> $OUT 2>&1 || { echo "plain run failed" >> $OUT; tail -20 $OUT; exit 1; }
EXTRA=""
[ "$TOOL" = "racecheck" ] && EXTRA="--racecheck-report all"
[ "$TOOL" = "initcheck" ] && EXTRA="--track-unused-memory no"
echo "# compute-sanitizer --tool $TOOL $EXTRA python tools/sanitize_target.py $ARGS" >> $OUT
timeout 1400 compute-sanitizer --tool $TOOL $EXTRA --print-limit 40 python tools/sanitize_target.py $ARGS >> $OUT 2>&1
echo "# exit code $?" >> $OUT
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE TARGET|Error|hazard|# exit" $OUT | tail -30
