#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/t4.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t4.log
timeout 900 python profiles/scratch/e2e_timer.py > gpurun_out/e2e_timer.log 2>&1; echo "rc=$?"
