#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/t3.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t3.log
timeout 900 python bench.py --cells 128 128 128 --no-cpu-baseline --breakdown > gpurun_out/b3_128.json 2> gpurun_out/b3_128.log; echo "bench rc=$?"
grep -E "push\+deposit|TOTAL|sort species|sync_particles" gpurun_out/b3_128.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b3_128.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
e=d['e2e']; print({k:e[k] for k in ('value','ms_per_step','h2d_seconds','d2h_seconds','step_ms')})
PY
