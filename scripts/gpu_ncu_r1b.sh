#!/bin/bash
# ncu --set full captures of the current kernels (one GPU): SA passes at the pipeline's 84-CTA limit and at the
# full grid, the tcgen05 rollout, the decoder epilogue; plus the launch list of a short pipelined bench run
mkdir -p gpurun_out
SA_CTAS=84 REPS=2 timeout 500 ncu --set full --clock-control none --import-source on -k regex:sa_pass_kernel -s 2 -c 2 -o gpurun_out/prof_sa_pass_84 -f python scripts/run_hot_once.py > gpurun_out/ncu_sa84.log 2>&1; echo "sa84 rc=$?"
REPS=2 timeout 500 ncu --set full --clock-control none --import-source on -k regex:sa_pass_kernel -s 2 -c 2 -o gpurun_out/prof_sa_pass_148 -f python scripts/run_hot_once.py > gpurun_out/ncu_sa148.log 2>&1; echo "sa148 rc=$?"
REPS=2 timeout 500 ncu --set full --clock-control none --import-source on -k regex:ro_umma_forward -s 1 -c 1 -o gpurun_out/prof_ro_r1b -f python scripts/run_hot_once.py > gpurun_out/ncu_ro.log 2>&1; echo "ro rc=$?"
REPS=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:"decode_combine|seg_argmax" -c 2 -o gpurun_out/prof_decode -f python scripts/run_hot_once.py > gpurun_out/ncu_dec.log 2>&1; echo "dec rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 60 -c 12 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1; echo "launchlist rc=$?"
ls -la gpurun_out/*.ncu-rep
