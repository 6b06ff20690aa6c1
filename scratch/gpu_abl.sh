#!/bin/bash
mkdir -p gpurun_out
DNS_SLAM_B200_LIB=$PWD/dns_slam_b200/libdns_slam_b200_ablate.so timeout 300 python scratch/abl.py > gpurun_out/abl.log 2>&1; echo "rc=$?"; tail -10 gpurun_out/abl.log
