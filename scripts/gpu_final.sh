#!/bin/bash
# evidence run: tests, smoke, bench (both arms), launch list, ncu --set full of the main kernels
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 600 python -m pytest tests -q -m gpu --tb=short > gpurun_out/t_all.log 2>&1; echo "tests rc=$?" >> gpurun_out/rc.txt
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.txt
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?" >> gpurun_out/rc.txt
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?" >> gpurun_out/rc.txt
if [ "$1" = "ncu" ]; then
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 60 -c 12 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1; echo "launchlist rc=$?" >> gpurun_out/rc.txt
SA_CTAS=84 REPS=2 timeout 500 ncu --set full --clock-control none --import-source on -k regex:sa_pass -s 2 -c 2 -o gpurun_out/prof_sa_pass_84 -f python scripts/run_hot_once.py > gpurun_out/ncu_sa84.log 2>&1; echo "ncu sa84 rc=$?" >> gpurun_out/rc.txt
REPS=2 timeout 500 ncu --set full --clock-control none --import-source on -k regex:sa_pass -s 2 -c 2 -o gpurun_out/prof_sa_pass_148 -f python scripts/run_hot_once.py > gpurun_out/ncu_sa148.log 2>&1; echo "ncu sa148 rc=$?" >> gpurun_out/rc.txt
REPS=2 timeout 500 ncu --set full --clock-control none --import-source on -k regex:sa_update_kernel -s 4 -c 1 -o gpurun_out/prof_sa_update -f python scripts/run_hot_once.py > gpurun_out/ncu_upd.log 2>&1; echo "ncu upd rc=$?" >> gpurun_out/rc.txt
REPS=2 timeout 500 ncu --set full --clock-control none --import-source on -k regex:ro_umma_forward -s 1 -c 1 -o gpurun_out/prof_ro -f python scripts/run_hot_once.py > gpurun_out/ncu_ro.log 2>&1; echo "ncu ro rc=$?" >> gpurun_out/rc.txt
REPS=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:"decode_combine|seg_argmax" -c 2 -o gpurun_out/prof_decode -f python scripts/run_hot_once.py > gpurun_out/ncu_dec.log 2>&1; echo "ncu dec rc=$?" >> gpurun_out/rc.txt
fi
cat gpurun_out/rc.txt; tail -3 gpurun_out/t_all.log; tail -1 gpurun_out/smoke.log; cat gpurun_out/bench_ref.json; cat gpurun_out/bench_full.json
