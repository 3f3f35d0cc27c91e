#!/bin/bash
for m in off on off on; do
  if [ $m = on ]; then export LPIC_REC_PREFETCH=1; else unset LPIC_REC_PREFETCH; fi
  echo "== prefetch $m"
  timeout 600 python bench.py --cells 128 128 128 --steps 4 --warmup 22 --no-e2e --no-cpu-baseline --breakdown 2>&1 >/dev/null | grep -E "push\+deposit|TOTAL"
done
