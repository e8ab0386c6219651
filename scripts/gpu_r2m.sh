#!/bin/bash
# job M: rollout with q | k | v over X ‖ Y (deeper weight ring at d = 256): parity, per-config times, phases
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_wrappers.py tests/test_gpu_training.py -q -m gpu --tb=short -x -k "rollout or ro_ or Rollout or slotformer or SlotFormer or train" > gpurun_out/t_ro.log 2>&1; echo "ro tests rc=$?"; tail -5 gpurun_out/t_ro.log
SKIP_SA=1 timeout 300 python scripts/prof_configs.py > gpurun_out/prof_configs.txt 2>&1; cat gpurun_out/prof_configs.txt
timeout 100 python scripts/prof_ro_phases.py 0 > gpurun_out/ro_phases_cfg2.txt 2>&1; cat gpurun_out/ro_phases_cfg2.txt
RO_CASE=ro_cfg3 RO_B=32 timeout 100 python scripts/prof_ro_phases.py 0 > gpurun_out/ro_phases_cfg3.txt 2>&1; cat gpurun_out/ro_phases_cfg3.txt
