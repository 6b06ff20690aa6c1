#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_loops.py -q --timeout 200 > gpurun_out/loops.log 2>&1; echo "rc=$?"; grep -n "^E   \|FAILED\|passed\|failed\|Error" gpurun_out/loops.log | head -30; tail -5 gpurun_out/loops.log
