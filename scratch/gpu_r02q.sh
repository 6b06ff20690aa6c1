#!/bin/bash
# GPU suite, smoke, our arm (defaults), then the ncu evidence of the same build (the reference arm is unchanged: r02p)
mkdir -p gpurun_out
TAG=${1:-r02q}
timeout 900 python -m pytest tests -x -m gpu -q --timeout 150 > gpurun_out/${TAG}_all.log 2>&1
echo "all rc=$?"; grep -n "AssertionError\|^E   .*assert\|FAILED\|passed\|failed\|Timeout" gpurun_out/${TAG}_all.log | head -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
b=json.load(open('gpurun_out/${TAG}_bench.json'))
print(b['value'], b['ms_per_step'], b['e2e'], b['roofline']['frac'], b['roofline']['traffic'], {k: round(v,3) for k,v in b['roofline']['phase_ms_per_step'].items()})
print(b['cpu_baseline'])
PY
timeout 900 bash profiles/run_ncu.sh $TAG 2>&1 | tail -3
