#!/bin/bash
# round-2 GPU job B: whole GPU suite, rollout errors, bench (both arms), launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --tb=short > gpurun_out/t_all.log 2>&1; echo "all tests rc=$?" > gpurun_out/rc.txt
timeout 200 python scripts/ro_errors.py > gpurun_out/ro_errors.txt 2>&1
timeout 900 python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo "bench rc=$?" >> gpurun_out/rc.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r2.json 2> gpurun_out/bench_ref_r2.err; echo "bench ref rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -15 gpurun_out/t_all.log; cat gpurun_out/ro_errors.txt; cat gpurun_out/bench_r2.json; tail -3 gpurun_out/bench_r2.err; cat gpurun_out/bench_ref_r2.json
