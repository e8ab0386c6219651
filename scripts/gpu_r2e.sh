#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/prof_sa_tc.py > gpurun_out/prof_sa_tc.txt 2>&1; echo "prof_sa_tc rc=$?"
head -100 gpurun_out/prof_sa_tc.txt
