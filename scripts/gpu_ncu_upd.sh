#!/bin/bash
mkdir -p gpurun_out
CHUNKS=384 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sa_update_kernel -s 1 -c 1 -o gpurun_out/prof_upd_r1 -f python scripts/prof_sa.py > gpurun_out/ncu_upd.log 2>&1; echo "ncu rc=$?"
