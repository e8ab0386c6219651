#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 3 4 7; do SFB_DBG=$d CHUNKS=384 timeout 120 python scripts/prof_sa.py; done > gpurun_out/sa_variants.txt 2>&1
cat gpurun_out/sa_variants.txt
CHUNKS=384 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sa_pass_kernel -s 2 -c 2 -o gpurun_out/prof_pass_r1 -f python scripts/prof_sa.py > gpurun_out/ncu_pass.log 2>&1; echo "ncu rc=$?"
