#!/bin/bash
mkdir -p gpurun_out
DNS_SLAM_B200_LIB=$PWD/dns_slam_b200/libdns_slam_b200_ablate.so timeout 300 python scratch/jimg_ab.py > gpurun_out/jimg_ab.log 2>&1; echo "rc=$?"; head -3 gpurun_out/jimg_ab.log | cut -c1-300
timeout 800 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/jimg_tests.log 2>&1; echo "tests rc=$?"; grep -n "^E   \|FAILED\|passed\|failed" gpurun_out/jimg_tests.log | head -20
