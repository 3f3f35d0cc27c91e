#!/bin/bash
# A/B of the record-layout variant of k_push_tile (LPIC_TILE_AOS=1: SoA -> records -> kernel -> SoA) at a decayed slot order
set -x
mkdir -p gpurun_out
LPIC_TILE_AOS=1 timeout 600 python -m pytest tests/test_gpu_bench_shapes.py -x -q -m gpu > gpurun_out/aos_parity.log 2>&1; echo "parity rc=$?"
tail -3 gpurun_out/aos_parity.log
for m in soa aos; do
  if [ $m = aos ]; then export LPIC_TILE_AOS=1; else unset LPIC_TILE_AOS; fi
  timeout 600 python bench.py --cells 128 128 128 --steps 4 --warmup 22 --no-e2e --no-cpu-baseline --breakdown > gpurun_out/aos_bench_$m.log 2>&1
  grep -E "push\+deposit|TOTAL|sort species" gpurun_out/aos_bench_$m.log
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_push_tile|k_to_rec|k_from_rec|k_tile_perm|k_list_particles' --csv --log-file gpurun_out/aos_launches_$m.csv python bench.py --cells 128 128 128 --steps 2 --warmup 24 --no-e2e --no-cpu-baseline > gpurun_out/aos_ncu_$m.log 2>&1
  tail -12 gpurun_out/aos_launches_$m.csv | cut -c1-200
done
