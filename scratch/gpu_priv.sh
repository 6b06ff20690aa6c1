#!/bin/bash
mkdir -p gpurun_out
DNS_SLAM_B200_LIB=$PWD/dns_slam_b200/libdns_slam_b200_ablate.so timeout 300 python scratch/priv_ab.py > gpurun_out/priv_ab.log 2>&1; echo "ab rc=$?"; cat gpurun_out/priv_ab.log | tail -15
timeout 300 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity_configs.py tests/test_gpu_framestep.py -q --timeout 100 > gpurun_out/priv_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/priv_tests.log
