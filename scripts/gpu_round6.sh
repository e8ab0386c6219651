#!/bin/bash
mkdir -p gpurun_out
CHUNKS=384 timeout 100 python scripts/prof_sa.py > gpurun_out/sa_times.txt 2>&1; cat gpurun_out/sa_times.txt
timeout 100 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "slot_attention_vs_reference" --tb=short 2>&1 | tail -3
CHUNKS=384 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sa_pass_kernel -s 7 -c 1 -o gpurun_out/prof_pass2_r1 -f python scripts/prof_sa.py > gpurun_out/ncu_pass2.log 2>&1; echo "ncu rc=$?"
