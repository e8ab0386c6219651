#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
timeout 120 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "rollout" --tb=short -x > gpurun_out/t_ro.log 2>&1; echo "ro rc=$?" >> gpurun_out/rc.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -15 gpurun_out/t_ro.log; python -c "
import json; d=json.load(open('gpurun_out/bench3.json')); print('ms/step',d['ms_per_step'],'sa ms',d['roofline']['ms_per_launch'],'ro ms',d['roofline_rollout']['ms_per_launch'], 'value', d['value'])"; tail -3 gpurun_out/bench3.err
