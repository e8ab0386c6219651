#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "slot_attention or empty" --tb=short > gpurun_out/t_sa.log 2>&1; echo "sa rc=$?" >> gpurun_out/rc.txt
CHUNKS=384 timeout 100 python scripts/prof_sa.py > gpurun_out/sa_times.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 7 --csv --log-file gpurun_out/launches5.csv python scripts/prof_sa.py > /dev/null 2>&1
cat gpurun_out/rc.txt; tail -12 gpurun_out/t_sa.log; cat gpurun_out/sa_times.txt
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches5.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows: print(r[4][:50], r[-1])
PY
