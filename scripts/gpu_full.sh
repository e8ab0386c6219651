#!/bin/bash
# full GPU validation: all gpu tests, smoke, bench, reference arm
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 600 python -m pytest tests -q -m gpu --tb=short > gpurun_out/t_all.log 2>&1; echo "tests rc=$?" >> gpurun_out/rc.txt
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.txt
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?" >> gpurun_out/rc.txt
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -8 gpurun_out/t_all.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err; cat gpurun_out/bench_ref.json
