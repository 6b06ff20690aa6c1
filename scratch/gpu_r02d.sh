#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -q -k formats > gpurun_out/r02d_fmt.log 2>&1; tail -15 gpurun_out/r02d_fmt.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02d_all.log 2>&1
echo "all rc=$?"; grep -n "AssertionError\|^E   .*assert\|FAILED\|passed\|failed" gpurun_out/r02d_all.log | head -40
