#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "slot_attention or empty or rejects" --tb=short > gpurun_out/t_sa.log 2>&1; echo "sa rc=$?" >> gpurun_out/rc.txt
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.txt
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench rc=$?" >> gpurun_out/rc.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/launches2.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
cat gpurun_out/rc.txt; tail -15 gpurun_out/t_sa.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench2.json; tail -5 gpurun_out/bench2.err
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches2.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows[:16]: print(r[4][:60], r[-1], r[-2])
PY
