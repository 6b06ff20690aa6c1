#!/bin/bash
# L1 / occupancy trade-off of the two point kernels re-measured after the code-size pass (ablate build: DNS_FWD_CARVE / DNS_BWD_CARVE)
mkdir -p gpurun_out; : > gpurun_out/carve.log
lib=$PWD/dns_slam_b200/libdns_slam_b200_ablate.so
for cfg in "72 86" "58 86" "86 86" "72 72" "72 100"; do
  set -- $cfg
  echo "== fwd carve $1 bwd carve $2" >> gpurun_out/carve.log
  DNS_FWD_CARVE=$1 DNS_BWD_CARVE=$2 DNS_SLAM_B200_LIB=$lib timeout 100 python scratch/time_core.py 2>&1 | tail -1 | cut -c1-140 >> gpurun_out/carve.log
done
cat gpurun_out/carve.log
