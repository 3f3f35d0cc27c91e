#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/t7.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t7.log
timeout 900 python bench.py --cells 128 128 128 --no-cpu-baseline > gpurun_out/b7_128.json 2> gpurun_out/b7_128.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b7_128.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
e=d['e2e']; print({k:e[k] for k in ('value','ms_per_step','h2d_seconds','d2h_seconds','h2d_bytes_per_step','d2h_bytes_per_step')})
PY
timeout 600 python examples/lwfa.py > gpurun_out/f_lwfa2.log 2>&1; echo "lwfa rc=$?"; tail -2 gpurun_out/f_lwfa2.log
