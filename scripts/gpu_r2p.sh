#!/bin/bash
# job P: compute-sanitizer over the slot transition kernel + ncu capture of it
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  SAN_ONLY=transition timeout 400 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_cases.py > gpurun_out/sanitize_tr_$tool.log 2>&1
  echo "$tool rc=$?"; tail -4 gpurun_out/sanitize_tr_$tool.log
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:transition_kernel -s 3 -c 2 -o gpurun_out/prof_transition -f python scripts/prof_transition.py 4 148 > gpurun_out/ncu_tr.log 2>&1; echo "ncu rc=$?"
