#!/bin/bash
# round-2 GPU job C: f1 route timing, per-role timeline of the tcgen05 SA passes
mkdir -p gpurun_out
timeout 300 python scripts/time_f1.py > gpurun_out/time_f1.txt 2>&1; echo "time_f1 rc=$?"
timeout 300 python scripts/prof_sa_tc.py > gpurun_out/prof_sa_tc.txt 2>&1; echo "prof_sa_tc rc=$?"
cat gpurun_out/time_f1.txt; head -150 gpurun_out/prof_sa_tc.txt
