#!/bin/bash
# ncu evidence for bench.py (one GPU).  Usage on the GPU box:  bash profiles/run_ncu.sh <tag>
# 1) plain run must exit 0; 2) launch list (gpu__time_duration); 3) --set full of the top kernels.
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 1 --rays-per-gpu 32768 --no-extra --no-cpu"
$CMD > $OUT/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_ray_tc2|k_point_fwd_tc2|k_point_bwd_tc2|k_dw_img|k_featmerge' -s 12 -c 12 -o $OUT/prof_$TAG $CMD > $OUT/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
ls -la $OUT | tail -8
