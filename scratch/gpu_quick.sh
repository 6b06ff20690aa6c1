#!/bin/bash
# quick, bounded: the fused-path tests with a per-test timeout
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity_configs.py tests/test_gpu_inference.py tests/test_gpu_edges.py -q --timeout 60 > gpurun_out/quick.log 2>&1
echo "rc=$?"; grep -n "AssertionError\|^E   \|FAILED\|passed\|failed\|Timeout" gpurun_out/quick.log | head -20
timeout 300 python -m pytest tests/test_gpu_pipeline.py -q --timeout 100 -x > gpurun_out/quick2.log 2>&1
echo "rc=$?"; grep -n "AssertionError\|^E   \|FAILED\|passed\|failed\|Timeout" gpurun_out/quick2.log | head -20; tail -5 gpurun_out/quick2.log
