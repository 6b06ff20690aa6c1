#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -m gpu -q --timeout 150 > gpurun_out/micro_tests.log 2>&1; echo "tests rc=$?"; grep -n "^E   \|FAILED\|passed\|failed" gpurun_out/micro_tests.log | head
timeout 300 python bench.py --no-extra --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], {k: round(v,3) for k,v in d['roofline']['phase_ms_per_step'].items()})"
timeout 200 python scratch/cfg1_time.py 2>&1 | cut -c1-140
