#!/bin/bash
# round-2 GPU job D: tcgen05 SA passes with eager aggregation + dense-row slot update: correctness, timing, per-kernel profile
mkdir -p gpurun_out
timeout 120 python scripts/check_sa_tc.py > gpurun_out/check_sa_tc.log 2>&1; rc=$?; echo "check rc=$rc"
cat gpurun_out/check_sa_tc.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_encoder_tail.py tests/test_wrappers.py -q -m gpu --tb=short -x > gpurun_out/t_sa.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/t_sa.log
sed -i 's/timeout 300 ncu/timeout 100 ncu/' scripts/gpu_sa_profile.sh
bash scripts/gpu_sa_profile.sh > gpurun_out/sa_profile.txt 2>&1
grep -A7 "tc_0.csv\|tc_84.csv" gpurun_out/sa_profile.txt
