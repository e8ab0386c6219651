#!/bin/bash
# job J: rollout after the LayerNorm rewrite: parity, per-phase profile, pipeline
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short -x -k "rollout" > gpurun_out/t_ro.log 2>&1; echo "ro tests rc=$?"; tail -3 gpurun_out/t_ro.log
timeout 100 python scripts/ro_errors.py > gpurun_out/ro_errors.txt 2>&1; cat gpurun_out/ro_errors.txt
timeout 100 python scripts/prof_ro_phases.py 0 > gpurun_out/ro_phases_r2.txt 2>&1; cat gpurun_out/ro_phases_r2.txt
timeout 100 python scripts/ab_pipeline.py 0 > gpurun_out/ab_pipeline3.txt 2>&1; cat gpurun_out/ab_pipeline3.txt
