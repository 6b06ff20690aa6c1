#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scratch/tv_one.py scannet 3 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/tv_launches.csv python scratch/tv_one.py scannet 3 > gpurun_out/tv_ncu.log 2>&1
tail -12 gpurun_out/tv_launches.csv | cut -d, -f5,10- | cut -c1-160
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_point_bwd_tc2 -s 2 -c 1 -o gpurun_out/tv_bwd python scratch/tv_one.py scannet 3 > gpurun_out/tv_ncu2.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_point_fwd_tc2 -s 2 -c 1 -o gpurun_out/tv_fwd python scratch/tv_one.py scannet 3 > gpurun_out/tv_ncu3.log 2>&1; echo rc=$?
