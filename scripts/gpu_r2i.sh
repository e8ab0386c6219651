#!/bin/bash
# job I: NaN fix under memcheck, 512-pixel items timing, SA/encoder tests
mkdir -p gpurun_out
RO_B=64 timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_cases.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -v "^=========     " gpurun_out/sanitize_memcheck.log | tail -16
timeout 120 python scripts/check_sa_tc.py > gpurun_out/check_sa_tc.log 2>&1; echo "check rc=$?"; tail -5 gpurun_out/check_sa_tc.log
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_encoder_tail.py tests/test_wrappers.py -q -m gpu --tb=short -x > gpurun_out/t_sa.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_sa.log
timeout 100 python scripts/ab_pipeline.py 0 > gpurun_out/ab_pipeline2.txt 2>&1; cat gpurun_out/ab_pipeline2.txt
