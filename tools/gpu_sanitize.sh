#!/bin/bash
# compute-sanitizer over every kernel of the library; logs into gpurun_out/
TAG=${1:-r02}
mkdir -p gpurun_out
python tools/sanitize_target.py > gpurun_out/${TAG}_sanitize_plain.log 2>&1; echo "plain rc=$?"; tail -1 gpurun_out/${TAG}_sanitize_plain.log
for TOOL in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $TOOL --print-limit 20 python tools/sanitize_target.py > gpurun_out/${TAG}_sanitizer_$TOOL.log 2>&1
  echo "$TOOL rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target ok" gpurun_out/${TAG}_sanitizer_$TOOL.log | tail -3
done
