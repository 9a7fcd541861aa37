#!/bin/bash
# compute-sanitizer over every kernel family at smoke sizes.  ONE tool per gpurun call (B200_PROFILING.md):
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh memcheck'      (also: racecheck | synccheck | initcheck)
# Logs: gpurun_out/r02_sanitizer_<tool>.txt (copied to profiles/).
TOOL=${1:-memcheck}
mkdir -p gpurun_out
if [ "$TOOL" = "checked" ]; then
  # compute-sanitizer is closed on this pool: the checked build (make checked) asserts the kernels' invariants instead
  OUT=gpurun_out/r02_checked_build.txt
  echo "# B200SORT_LIB=libb200sort_checked.so python tools/sanitize_target.py big  (library built by: make checked)" > $OUT
  B200SORT_LIB=libb200sort_checked.so timeout 900 python tools/sanitize_target.py big >> $OUT 2>&1
  echo "# exit code $?" >> $OUT
  tail -8 $OUT
  exit 0
fi
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
