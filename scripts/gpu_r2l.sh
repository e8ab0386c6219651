#!/bin/bash
# job L: per-config operator times and the rollout's per-phase budget at HEAD
mkdir -p gpurun_out
timeout 300 python scripts/prof_configs.py > gpurun_out/prof_configs.txt 2>&1; cat gpurun_out/prof_configs.txt
timeout 100 python scripts/prof_ro_phases.py 0 > gpurun_out/ro_phases_cfg2.txt 2>&1; cat gpurun_out/ro_phases_cfg2.txt
RO_CASE=ro_cfg3 RO_B=32 timeout 100 python scripts/prof_ro_phases.py 0 > gpurun_out/ro_phases_cfg3.txt 2>&1; cat gpurun_out/ro_phases_cfg3.txt
