#!/bin/bash
mkdir -p gpurun_out
DNS_SLAM_B200_LIB=$PWD/dns_slam_b200/libdns_slam_b200_ablate.so timeout 300 python scratch/tv_ab.py > gpurun_out/tv_ab.log 2>&1; echo "rc=$?"; grep "agg 1.0\|agg 0 " gpurun_out/tv_ab.log | cut -c1-150
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_ops.py tests/test_gpu_loops.py tests/test_gpu_framestep.py tests/test_gpu_edges.py tests/test_gpu_parity_configs.py tests/test_gpu_fuzz.py tests/test_gpu_tc.py -q -m gpu --timeout 200 > gpurun_out/tv_tests.log 2>&1; echo "tests rc=$?"; grep -n "^E   \|FAILED\|passed\|failed" gpurun_out/tv_tests.log | head
for v in base; do DNS_SLAM_B200_LIB=$PWD/dns_slam_b200/libdns_slam_b200.so timeout 200 python scratch/time_core.py 2>&1 | tail -2 | cut -c1-200; done
timeout 200 python scratch/cfg1_time.py 2>&1 | cut -c1-200
