#!/bin/bash
# final single-GPU measurement batch of round 2 (record layout)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/f_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/f_tests.log
timeout 1500 python bench.py --breakdown > gpurun_out/f_bench_default.json 2> gpurun_out/f_bench_default.log; echo "default bench rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_push_tile --launch-skip 20 --launch-count 2 -o gpurun_out/f_prof_tile python bench.py --cells 128 128 128 --steps 2 --warmup 10 --no-e2e --no-cpu-baseline > gpurun_out/f_prof_tile.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f_launches.csv python bench.py --cells 128 128 128 --steps 2 --warmup 24 --no-e2e --no-cpu-baseline > gpurun_out/f_launches.log 2>&1; echo "ncu list rc=$?"
timeout 900 python bench.py --scaling strong --no-e2e --no-cpu-baseline --breakdown > gpurun_out/f_bench_strong.json 2> gpurun_out/f_bench_strong.log; echo "strong rc=$?"
timeout 900 python examples/laser_target_3d.py > gpurun_out/f_laser3d.log 2>&1; echo "laser3d rc=$?"; tail -2 gpurun_out/f_laser3d.log
timeout 600 python examples/lwfa.py > gpurun_out/f_lwfa.log 2>&1; echo "lwfa rc=$?"; tail -2 gpurun_out/f_lwfa.log
