#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "rollout" --tb=short > gpurun_out/t_ro.log 2>&1; echo "ro rc=$?" >> gpurun_out/rc.txt
timeout 100 python scripts/prof_ro.py > gpurun_out/ro_timeline.txt 2>&1
cat gpurun_out/rc.txt; tail -12 gpurun_out/t_ro.log; cat gpurun_out/ro_timeline.txt
