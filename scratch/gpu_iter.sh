#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_loops.py tests/test_gpu_framestep.py tests/test_gpu_pipeline.py -q --timeout 200 -x > gpurun_out/loops.log 2>&1; echo "rc=$?"; grep -n "^E   \|FAILED\|passed\|failed\|Error" gpurun_out/loops.log | head -30
timeout 400 python scratch/iter_time2.py > gpurun_out/iter_time2.log 2>&1; echo "rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/iter_time2.log'):
    if '{' in l:
        d=json.loads(l[l.index('{'):]); print(l.split()[0], {k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k!='note'})
    else: print(l[:300])
PY
