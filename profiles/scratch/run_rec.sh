#!/bin/bash
# full GPU test suite with the record layout, then A/B of the two layouts at 128^3 (decayed slot order, steps 23-26)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/rec_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/rec_tests.log
for m in rec soa; do
  if [ $m = soa ]; then export LPIC_PARTICLE_LAYOUT=soa; else unset LPIC_PARTICLE_LAYOUT; fi
  timeout 600 python bench.py --cells 128 128 128 --steps 4 --warmup 22 --no-e2e --no-cpu-baseline --breakdown > gpurun_out/rec_bench_$m.log 2>&1
  echo "== $m"; grep -E "push\+deposit|TOTAL|sort species|sync_particles" gpurun_out/rec_bench_$m.log; tail -1 gpurun_out/rec_bench_$m.log | cut -c1-200
done
