#!/bin/bash
# the whole GPU suite (as the driver runs it) + smoke + the default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu --timeout 300 > gpurun_out/full_tests.log 2>&1; echo "tests rc=$?"
grep -n "^E   \|FAILED\|passed\|failed\|Timeout" gpurun_out/full_tests.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/full_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel'])
print({k: round(v, 3) for k, v in d['roofline']['phase_ms_per_step'].items()})
e = d['extra']
for k in ('config1_tracking_1024x96', 'config2_mapping_4096x47_adam', 'tv_smoothness_63^3'): print(k, e[k])
print('iteration_scannet', {k: v for k, v in e['iteration_scannet'].items() if 'ms' in k})
print('config3', e['config3_scannet_200']['wall_s'], e['config3_scannet_200']['timings'])
PY
