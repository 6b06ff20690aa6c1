#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scratch/parity_table.py > gpurun_out/r02f_parity.log 2>&1; grep -v "table per level" gpurun_out/r02f_parity.log | tail -16
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02f_all.log 2>&1
echo "all rc=$?"; grep -n "AssertionError\|^E   .*assert\|FAILED\|passed\|failed" gpurun_out/r02f_all.log | head -40
timeout 600 python bench.py --steps 5 --warmup 2 --no-extra --no-cpu > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
b=json.load(open('gpurun_out/r02f_bench.json'))
print(b['value'], b['ms_per_step'], b['e2e']['value'], b['roofline']['phase_ms_per_step'])
PY
