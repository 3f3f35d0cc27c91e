#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bench_shapes.py tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -x -q -m gpu > gpurun_out/t9.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t9.log
for m in pairs nopairs pairs nopairs; do
  if [ $m = nopairs ]; then export LPIC_TILE_NO_PAIRS=1; else unset LPIC_TILE_NO_PAIRS; fi
  echo "== $m"
  timeout 600 python bench.py --cells 128 128 128 --steps 4 --warmup 22 --no-e2e --no-cpu-baseline --breakdown 2>&1 >/dev/null | grep -E "push\+deposit|TOTAL"
done
