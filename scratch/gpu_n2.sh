#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02k_n2.json 2> gpurun_out/r02k_n2.err
echo "n2 rc=$?"; tail -5 gpurun_out/r02k_n2.err
python - <<'PY'
import json
for l in open('gpurun_out/r02k_n2.json'):
    if l.startswith('{'):
        b=json.loads(l); print('N=2', b['value'], b['ms_per_step'], b['e2e']['value'], b['losses_last_step'][:9])
PY
timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu > gpurun_out/r02k_n1.json 2> gpurun_out/r02k_n1.err; echo "n1 rc=$?"
python - <<'PY'
import json
b=json.load(open('gpurun_out/r02k_n1.json')); print('N=1', b['value'], b['ms_per_step'], b['e2e']['value'], b['losses_last_step'][:9])
PY
