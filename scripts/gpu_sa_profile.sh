#!/bin/bash
# per-kernel durations + DRAM bytes of one Slot Attention call (bench workload): tcgen05 vs mma.sync passes, 148 / 84 CTAs
mkdir -p gpurun_out
for mode in tc mma; do for ctas in 0 84; do
  if [ $mode = mma ]; then export SA_NO_TC=1; else unset SA_NO_TC; fi
  SA_ONLY=1 SA_CTAS=$ctas REPS=3 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum --clock-control none -s 12 -c 6 --csv --log-file gpurun_out/sa_launches_${mode}_${ctas}.csv python scripts/run_hot_once.py > gpurun_out/sa_launches_${mode}_${ctas}.log 2>&1
  echo "$mode $ctas rc=$?"
done; done
python - <<'PY'
import csv, glob
for f in sorted(glob.glob('gpurun_out/sa_launches_*.csv')):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    hdr = rows[0]; ki, mi, vi = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value')
    ii = hdr.index('ID')
    by = {}
    for r in rows[1:]:
        by.setdefault((int(r[ii]), r[ki].split('(')[0][-60:]), {})[r[mi]] = r[vi]
    print('==', f)
    for (i, k), m in sorted(by.items()):
        print(f"{i:3d} {k:<62} {float(m['gpu__time_duration.sum'].replace(',',''))/1e3:8.1f} us  rd {float(m['dram__bytes_read.sum'].replace(',',''))/1e6:7.1f} MB  wr {float(m['dram__bytes_write.sum'].replace(',',''))/1e6:7.1f} MB  tensor {m.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','-')}%  inst {m.get('sm__inst_executed.sum','-')}")
PY
