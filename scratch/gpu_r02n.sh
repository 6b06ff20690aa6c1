#!/bin/bash
# the driver's sequence: GPU suite, smoke, reference arm, our arm (defaults), then the ncu evidence
mkdir -p gpurun_out
TAG=${1:-r02n}
timeout 900 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/${TAG}_all.log 2>&1
echo "all rc=$?"; grep -n "AssertionError\|^E   .*assert\|FAILED\|passed\|failed\|Timeout" gpurun_out/${TAG}_all.log | head -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err; echo "ref rc=$?"; tail -1 gpurun_out/${TAG}_ref.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
b=json.load(open('gpurun_out/${TAG}_bench.json'))
print(b['value'], b['ms_per_step'], b['e2e'], b['roofline']['frac'], b['roofline']['traffic'], {k: round(v,3) for k,v in b['roofline']['phase_ms_per_step'].items()})
print(b['cpu_baseline'])
r=json.load(open('gpurun_out/${TAG}_ref.json')); print('ref', r['value'], r['ms_per_step'])
PY
timeout 600 bash profiles/run_ncu.sh $TAG 2>&1 | tail -3
