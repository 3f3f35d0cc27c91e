#!/bin/bash
for nw in 8 12 16; do
  export LPIC_TILE_NW=$nw
  echo "== warps per CTA $nw"
  timeout 600 python bench.py --cells 128 128 128 --steps 4 --warmup 22 --no-e2e --no-cpu-baseline --breakdown 2>&1 >/dev/null | grep -E "push\+deposit|TOTAL"
done
