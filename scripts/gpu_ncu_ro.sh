#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ro_umma_forward -s 3 -c 1 -o gpurun_out/prof_ro_umma_r1 -f python scripts/prof_ro.py > gpurun_out/ncu_ro_umma.log 2>&1; echo "ncu rc=$?"
