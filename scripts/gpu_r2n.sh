#!/bin/bash
# job N: fused slot transition (f3): parity + timing
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_transition.py -q -m gpu --tb=short -x > gpurun_out/t_tr.log 2>&1; echo "tr tests rc=$?"; tail -15 gpurun_out/t_tr.log
timeout 100 python scripts/prof_transition.py 4 64 148 > gpurun_out/prof_transition.txt 2>&1; cat gpurun_out/prof_transition.txt
timeout 200 python scripts/time_transition.py > gpurun_out/time_transition.txt 2>&1; cat gpurun_out/time_transition.txt | tail -20
