#!/bin/bash
# ncu --set full of the tcgen05 SA passes at 84 CTAs (SM-bound regime)
mkdir -p gpurun_out
SA_ONLY=1 SA_CTAS=84 REPS=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:sa_pass_tc -s 2 -c 2 -o gpurun_out/prof_sa_tc_84 -f python scripts/run_hot_once.py > gpurun_out/ncu_sa_tc84.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_sa_tc84.log
