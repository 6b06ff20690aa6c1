#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02k_n8.json 2> gpurun_out/r02k_n8.err
echo "n2 rc=$?"; tail -5 gpurun_out/r02k_n8.err
python - <<'PY'
import json
for l in open('gpurun_out/r02k_n8.json'):
    if l.startswith('{'):
        b=json.loads(l); print('N=8', b['value'], b['ms_per_step'], b['e2e']['value'], b['losses_last_step'][:9])
PY
