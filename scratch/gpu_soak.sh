#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scratch/soak.py > gpurun_out/soak.log 2>&1; echo "rc=$?"; tail -14 gpurun_out/soak.log
timeout 300 compute-sanitizer --tool memcheck --launch-timeout 0 python -m pytest tests/test_gpu_framestep.py -q --timeout 280 -x -k "48-True or 300" > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?"; grep -n "ERROR SUMMARY\|Invalid\|passed\|failed" gpurun_out/sanitizer.log | head
