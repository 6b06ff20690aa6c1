#!/bin/bash
mkdir -p gpurun_out
DNS_SLAM_B200_LIB=$PWD/dns_slam_b200/libdns_slam_b200_ablate.so timeout 300 python scratch/tv_ab.py > gpurun_out/tv_ab.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/tv_ab.log
