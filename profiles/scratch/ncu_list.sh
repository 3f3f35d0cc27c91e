#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/rec_launches.csv python bench.py --cells 128 128 128 --steps 2 --warmup 24 --no-e2e --no-cpu-baseline > gpurun_out/rec_ncu.log 2>&1
echo rc=$?
