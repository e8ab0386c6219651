#!/bin/bash
mkdir -p gpurun_out
timeout 100 python scripts/prof_ro.py > gpurun_out/ro_timeline.txt 2>&1; cat gpurun_out/ro_timeline.txt
