#!/bin/bash
# job O: rollout A/B (ring depth), parity + pipeline timing
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short -x -k "rollout" > gpurun_out/t_ro.log 2>&1; echo "ro tests rc=$?"; tail -3 gpurun_out/t_ro.log
SKIP_SA=1 timeout 300 python scripts/prof_configs.py > gpurun_out/prof_configs.txt 2>&1; cat gpurun_out/prof_configs.txt
timeout 100 python scripts/ab_pipeline.py 0 > gpurun_out/ab_pipeline.txt 2>&1; cat gpurun_out/ab_pipeline.txt
