#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_wrappers.py -q -m gpu --tb=short -x > gpurun_out/t_sa.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_sa.log
timeout 120 python scripts/check_sa_tc.py > gpurun_out/check_sa_tc.log 2>&1; echo "check rc=$?"; tail -4 gpurun_out/check_sa_tc.log
timeout 100 python scripts/ab_pipeline.py 0 > gpurun_out/ab_pipeline4.txt 2>&1; cat gpurun_out/ab_pipeline4.txt
