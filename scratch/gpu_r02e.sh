#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02e_all.log 2>&1
echo "all rc=$?"; grep -n "AssertionError\|^E   .*assert\|FAILED\|passed\|failed" gpurun_out/r02e_all.log | head -40
