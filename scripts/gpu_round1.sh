#!/bin/bash
# first GPU pass: parity tests (each group under its own timeout), smoke, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "slot_attention" --tb=short -x > gpurun_out/t_sa.log 2>&1; echo "sa rc=$?" >> gpurun_out/rc.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "rollout" --tb=short > gpurun_out/t_ro.log 2>&1; echo "ro rc=$?" >> gpurun_out/rc.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "not rollout and not slot_attention" --tb=short > gpurun_out/t_misc.log 2>&1; echo "misc rc=$?" >> gpurun_out/rc.txt
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.txt
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -5 gpurun_out/t_sa.log; tail -5 gpurun_out/t_ro.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
