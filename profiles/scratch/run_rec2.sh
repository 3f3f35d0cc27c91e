#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/rec_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/rec_tests.log
timeout 600 python bench.py --cells 128 128 128 --steps 4 --warmup 22 --no-e2e --no-cpu-baseline --breakdown > gpurun_out/rec_bench_rec.log 2>&1
grep -E "push\+deposit|TOTAL|sort species|sync_particles" gpurun_out/rec_bench_rec.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_push_tile --launch-skip 20 --launch-count 2 -o gpurun_out/prof_rec python bench.py --cells 128 128 128 --steps 2 --warmup 10 --no-e2e --no-cpu-baseline > gpurun_out/prof_rec.log 2>&1
echo "ncu rc=$?"
