#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02n}
S0=$SECONDS; timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err; echo "ref rc=$? wall $((SECONDS-S0)) s"; tail -1 gpurun_out/${TAG}_ref.err
S0=$SECONDS; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$? wall $((SECONDS-S0)) s"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
b=json.load(open('gpurun_out/${TAG}_bench.json'))
print(b['value'], b['ms_per_step'], b['e2e'], b['roofline']['frac'], b['roofline']['traffic'], {k: round(v,3) for k,v in b['roofline']['phase_ms_per_step'].items()})
print(b['cpu_baseline'])
r=json.load(open('gpurun_out/${TAG}_ref.json')); print('ref', r['value'], r['ms_per_step'])
PY
