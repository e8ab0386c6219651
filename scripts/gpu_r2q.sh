#!/bin/bash
# job Q: tcgen05 first pass with two 16-pixel stages per LayerNorm warp: parity, timing alone / capped, pipeline
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_wrappers.py tests/test_encoder_tail.py -q -m gpu --tb=short -x > gpurun_out/t_sa.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_sa.log
timeout 200 python scripts/time_f1.py 2>&1 | grep "SA on" | cut -c1-200
timeout 100 python scripts/ab_pipeline.py 0 > gpurun_out/ab_pipeline.txt 2>&1; cat gpurun_out/ab_pipeline.txt
SA_CTAS=84 timeout 200 python scripts/prof_sa_tc.py 2>&1 | grep -E "===|LN warp" | cut -c1-900
