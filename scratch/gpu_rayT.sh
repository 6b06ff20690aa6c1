#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/rayT.log
lib=$PWD/dns_slam_b200/libdns_slam_b200_ablate.so
for t in 256 128 256 128; do
  echo "== DNS_RAY_T=$t" >> gpurun_out/rayT.log
  DNS_RAY_T=$t DNS_SLAM_B200_LIB=$lib timeout 200 python scratch/cfg1_time.py >> gpurun_out/rayT.log 2>&1 || echo "failed" >> gpurun_out/rayT.log
done
cut -c1-250 gpurun_out/rayT.log
