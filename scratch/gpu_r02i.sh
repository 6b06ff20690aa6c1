#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/r02i_all.log 2>&1
echo "all rc=$?"; grep -n "AssertionError\|^E   .*assert\|FAILED\|passed\|failed\|Timeout" gpurun_out/r02i_all.log | head -30
timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02i_bench.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/r02i_bench.json'))
print(b['value'], b['ms_per_step'], b['e2e']['value'], {k: round(v,3) for k,v in b['roofline']['phase_ms_per_step'].items()})
PY
