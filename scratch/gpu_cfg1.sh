#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scratch/cfg1_time.py 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_loops.py tests/test_gpu_fused.py tests/test_gpu_parity_configs.py -q --timeout 200 > gpurun_out/loops.log 2>&1; echo "rc=$?"; grep -n "^E   \|FAILED\|passed\|failed\|Error" gpurun_out/loops.log | head
