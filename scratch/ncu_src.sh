#!/bin/bash
# source-level capture of the three big kernels (one launch each) at 32768 rays
set -u
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 1 --rays-per-gpu 32768 --no-extra --cpu-rays 256"
$CMD > $OUT/plain_src.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/plain_src.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'k_ray_tc|k_point_fwd_tc2|k_point_bwd_tc2' -s 6 -c 3 -o $OUT/prof_src $CMD > $OUT/ncu_src.log 2>&1
echo "capture rc=$?"
ls -la $OUT | tail -5
