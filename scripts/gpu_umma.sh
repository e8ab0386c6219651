#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_umma.py -q -m gpu --tb=short -x > gpurun_out/t_umma.log 2>&1; echo "umma rc=$?"
tail -30 gpurun_out/t_umma.log
