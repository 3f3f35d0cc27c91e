#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/t5.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t5.log
for m in rec soa; do
  if [ $m = soa ]; then export LPIC_PARTICLE_LAYOUT=soa; else unset LPIC_PARTICLE_LAYOUT; fi
  echo "== 3D 128^3 step 26, $m"
  timeout 600 python bench.py --cells 128 128 128 --steps 4 --warmup 22 --no-e2e --no-cpu-baseline --breakdown 2>&1 >/dev/null | grep -E "push\+deposit|TOTAL|sort species|sync_particles"
  echo "== 2D 2048^2 step 5 / step 25, $m"
  timeout 600 python profiles/scratch/bench2d.py 4
  timeout 600 python profiles/scratch/bench2d.py 24
done
unset LPIC_PARTICLE_LAYOUT
timeout 900 python bench.py --cells 128 128 128 --no-cpu-baseline > gpurun_out/b5_128.json 2> gpurun_out/b5_128.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b5_128.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
e=d['e2e']; print({k:e[k] for k in ('value','ms_per_step','h2d_seconds','d2h_seconds','step_ms')})
PY
