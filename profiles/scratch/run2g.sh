#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nccl.py -x -q -m gpu > gpurun_out/t2g.log 2>&1; echo "nccl tests rc=$?"; tail -3 gpurun_out/t2g.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --cells 128 128 128 --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/b2g_rec.json 2> gpurun_out/b2g_rec.log; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b2g_rec.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('multirank_parity'))
e=d['e2e']; print({k:e.get(k) for k in ('value','ms_per_step','h2d_seconds','d2h_seconds','step_ms')})
PY
