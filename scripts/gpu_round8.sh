#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 150 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "rollout" --tb=short > gpurun_out/t_ro.log 2>&1; echo "ro(umma) rc=$?" >> gpurun_out/rc.txt
SFB_RO_ENGINE=mma timeout 150 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "rollout" --tb=line > gpurun_out/t_ro_mma.log 2>&1; echo "ro(mma) rc=$?" >> gpurun_out/rc.txt
timeout 100 python scripts/prof_ro.py > gpurun_out/ro_timeline.txt 2>&1
cat gpurun_out/rc.txt; tail -25 gpurun_out/t_ro.log; tail -4 gpurun_out/t_ro_mma.log; head -3 gpurun_out/ro_timeline.txt
