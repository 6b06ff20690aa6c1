#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_ops.py tests/test_gpu_featmerge.py tests/test_gpu_loops.py tests/test_gpu_framestep.py tests/test_gpu_edges.py tests/test_gpu_parity_configs.py tests/test_gpu_fuzz.py tests/test_gpu_tc.py tests/test_gpu_inference.py -q -m gpu --timeout 200 > gpurun_out/roll_tests.log 2>&1; echo "tests rc=$?"; grep -n "^E   \|FAILED\|passed\|failed" gpurun_out/roll_tests.log | head
timeout 200 python scratch/time_core.py 2>&1 | tail -2 | cut -c1-120
timeout 200 python scratch/cfg1_time.py 2>&1 | cut -c1-200
timeout 200 python scratch/tv_ab.py 2>&1 | grep "agg 1.0" | cut -c1-100
timeout 300 python bench.py --no-extra --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], {k: round(v,3) for k,v in d['roofline']['phase_ms_per_step'].items()})"
