#!/bin/bash
mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/ray_tests.log 2>&1; echo "tests rc=$?"; grep -n "^E   \|FAILED\|passed\|failed" gpurun_out/ray_tests.log | head -20
timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu > gpurun_out/fm_bench.json 2> gpurun_out/fm_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/fm_bench.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/fm_bench.json'))
print(b['value'], b['ms_per_step'], b['e2e']['value'], {k: round(v,3) for k,v in b['roofline']['phase_ms_per_step'].items()})
PY
