#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_loops.py tests/test_gpu_pipeline.py tests/test_gpu_inference.py -q > gpurun_out/r02j_loops.log 2>&1
echo "rc=$?"; grep -n "AssertionError\|^E   \|FAILED\|passed\|failed\|Error" gpurun_out/r02j_loops.log | head -30
( time python examples/synthetic_slam.py scannet 200 5 ) > gpurun_out/r02j_cfg3.log 2>&1; tail -6 gpurun_out/r02j_cfg3.log
