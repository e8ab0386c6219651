#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/ab_pipeline.py 8 0 8 0 > gpurun_out/ab_pipeline.txt 2>&1; echo "ab rc=$?"
cat gpurun_out/ab_pipeline.txt
