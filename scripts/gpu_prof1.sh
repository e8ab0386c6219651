#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/prof_timeline.py > gpurun_out/timeline1.txt 2>&1; echo "timeline rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sa_forward_kernel -s 2 -c 1 -o gpurun_out/prof_sa_r1 -f python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_sa.log 2>&1; echo "ncu sa rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ro_forward_kernel -s 2 -c 1 -o gpurun_out/prof_ro_r1 -f python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_ro.log 2>&1; echo "ncu ro rc=$?"
cat gpurun_out/timeline1.txt
ls -la gpurun_out
