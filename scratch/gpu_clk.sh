#!/bin/bash
mkdir -p gpurun_out
DNS_SLAM_B200_LIB=$PWD/dns_slam_b200/libdns_slam_b200_ablate.so timeout 300 python scratch/phase_clk_fm.py > gpurun_out/phase_clk_fm.log 2>&1; echo "rc=$?"; tail -20 gpurun_out/phase_clk_fm.log
