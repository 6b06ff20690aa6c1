#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}; TAG=${2:-r02n}
nvidia-smi --query-gpu=name --format=csv,noheader | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err
echo "rc=$?"; tail -5 gpurun_out/${TAG}_n$N.err
python - <<PY
import json
for l in open('gpurun_out/${TAG}_n$N.json'):
    if l.startswith('{'):
        b=json.loads(l); print('N=$N', b['value'], b['ms_per_step'], b['e2e']['value'], b['losses_last_step'][:9])
PY
