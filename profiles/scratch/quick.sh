#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bench_shapes.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/q_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/q_tests.log
for m in rec soa; do
  if [ $m = soa ]; then export LPIC_PARTICLE_LAYOUT=soa; else unset LPIC_PARTICLE_LAYOUT; fi
  timeout 600 python bench.py --cells 128 128 128 --steps 4 --warmup 22 --no-e2e --no-cpu-baseline --breakdown > gpurun_out/q_bench_$m.log 2>&1
  echo "== $m"; grep -E "push\+deposit|TOTAL|sort species|sync_particles" gpurun_out/q_bench_$m.log
done
