#!/bin/bash
# compute-sanitizer over every kernel family (memcheck, racecheck, synccheck); logs -> gpurun_out/sanitize_*.log
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  RO_B=160 timeout 500 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_cases.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; tail -4 gpurun_out/sanitize_$tool.log
done
