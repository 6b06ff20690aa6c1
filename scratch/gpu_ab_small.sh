#!/bin/bash
# usage: gpu_ab_small.sh <suffix> [...]: small-batch configs (scratch/cfg1_time.py) per library variant, release library last
mkdir -p gpurun_out; : > gpurun_out/ab_small.log
for v in "$@" base "$@" base; do
  if [ "$v" = base ]; then lib=$PWD/dns_slam_b200/libdns_slam_b200.so; else lib=$PWD/dns_slam_b200/libdns_slam_b200_$v.so; fi
  echo "== $v" >> gpurun_out/ab_small.log
  DNS_SLAM_B200_LIB=$lib timeout 200 python scratch/cfg1_time.py >> gpurun_out/ab_small.log 2>&1 || echo "$v failed" >> gpurun_out/ab_small.log
done
cut -c1-330 gpurun_out/ab_small.log
