#!/bin/bash
# full GPU suite + N=1 bench + ncu evidence of the current build (bounded)
mkdir -p gpurun_out
TAG=${1:-r02m}
timeout 900 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/${TAG}_all.log 2>&1
echo "all rc=$?"; grep -n "AssertionError\|^E   .*assert\|FAILED\|passed\|failed\|Timeout" gpurun_out/${TAG}_all.log | head -30
timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
b=json.load(open('gpurun_out/${TAG}_bench.json'))
print(b['value'], b['ms_per_step'], b['e2e']['value'], {k: round(v,3) for k,v in b['roofline']['phase_ms_per_step'].items()})
PY
timeout 600 bash profiles/run_ncu.sh $TAG 2>&1 | tail -5
