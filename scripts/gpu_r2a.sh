#!/bin/bash
# round-2 combined GPU job A: SA per-kernel profile, new GPU tests, whole GPU suite
mkdir -p gpurun_out
bash scripts/gpu_sa_profile.sh > gpurun_out/sa_profile.txt 2>&1
timeout 600 python -m pytest tests/test_encoder_tail.py tests/test_gpu_training.py tests/test_wrappers.py tests/test_decode.py -q -m gpu --tb=short -x > gpurun_out/t_new.log 2>&1; echo "new tests rc=$?" > gpurun_out/rc.txt
timeout 900 python -m pytest tests -q -m gpu --tb=short > gpurun_out/t_all.log 2>&1; echo "all tests rc=$?" >> gpurun_out/rc.txt
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -25 gpurun_out/sa_profile.txt; tail -15 gpurun_out/t_new.log; tail -15 gpurun_out/t_all.log; tail -2 gpurun_out/smoke.log
