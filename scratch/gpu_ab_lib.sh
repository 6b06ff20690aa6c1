#!/bin/bash
# usage: gpu_ab_lib.sh <suffix> [<suffix> ...]   ('' = release library)
mkdir -p gpurun_out; : > gpurun_out/ab_lib.log
for v in "$@" base; do
  if [ "$v" = base ]; then lib=$PWD/dns_slam_b200/libdns_slam_b200.so; else lib=$PWD/dns_slam_b200/libdns_slam_b200_$v.so; fi
  DNS_SLAM_B200_LIB=$lib timeout 200 python scratch/time_core.py >> gpurun_out/ab_lib.log 2>&1 || echo "$v failed" >> gpurun_out/ab_lib.log
done
cat gpurun_out/ab_lib.log | tail -20
