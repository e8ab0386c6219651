#!/bin/bash
# round-2 GPU job H: whole suite, smoke, bench (ours)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --tb=short > gpurun_out/t_all.log 2>&1; echo "all tests rc=$?" > gpurun_out/rc.txt
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.txt
timeout 600 python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo "bench rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -8 gpurun_out/t_all.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench_r2.json; tail -5 gpurun_out/bench_r2.err
