#!/bin/bash
# L2 residency of the hash / gradient table against the streamed stash: access-policy window (ablate build, DNS_L2_WIN) and
# (the switches live in scratch/l2_residency_experiment.patch: `patch -p0 < scratch/l2_residency_experiment.patch`, then
#  `make ablate` and `make variant VARIANT=cs VFLAGS=-DDNS_STASH_CS`; measured slower, not part of the library)
# streaming stores of the tile images (variant build -DDNS_STASH_CS)
mkdir -p gpurun_out; : > gpurun_out/l2.log
A=$PWD/dns_slam_b200/libdns_slam_b200_ablate.so
echo "== ablate, no window" >> gpurun_out/l2.log;  DNS_L2_WIN=0 DNS_SLAM_B200_LIB=$A timeout 100 python scratch/time_core.py 2>&1 | tail -1 | cut -c1-130 >> gpurun_out/l2.log
echo "== ablate, L2 window" >> gpurun_out/l2.log;  DNS_L2_WIN=1 DNS_SLAM_B200_LIB=$A timeout 100 python scratch/time_core.py 2>&1 | tail -1 | cut -c1-130 >> gpurun_out/l2.log
echo "== streaming stash stores" >> gpurun_out/l2.log; DNS_SLAM_B200_LIB=$PWD/dns_slam_b200/libdns_slam_b200_cs.so timeout 100 python scratch/time_core.py 2>&1 | tail -1 | cut -c1-130 >> gpurun_out/l2.log
cat gpurun_out/l2.log
