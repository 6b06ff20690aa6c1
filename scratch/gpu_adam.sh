#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_adam.py tests/test_gpu_framestep.py tests/test_gpu_loops.py tests/test_gpu_pipeline.py -q -m gpu --timeout 200 > gpurun_out/adam_tests.log 2>&1; echo "tests rc=$?"; grep -n "^E   \|FAILED\|passed\|failed" gpurun_out/adam_tests.log | head
timeout 200 python scratch/cfg1_time.py 2>&1 | cut -c1-260 | tail -1
