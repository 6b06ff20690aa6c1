#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scratch/time_core.py > gpurun_out/img_time.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/img_time.log | cut -c1-200
timeout 800 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/img_tests.log 2>&1; echo "tests rc=$?"; grep -n "^E   \|FAILED\|passed\|failed" gpurun_out/img_tests.log | head -20
