#!/bin/bash
# ncu evidence for the ResNet stem kernels (one GPU).  Usage on the GPU box:  bash profiles/run_ncu_stem.sh <tag>
set -u
TAG=${1:-r01l}
OUT=gpurun_out
mkdir -p $OUT
CMD="python scratch/stem_time.py"
$CMD > $OUT/plain_stem_$TAG.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/plain_stem_$TAG.log; exit 1; }
cat $OUT/plain_stem_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:'k_stem' -s 8 -c 4 -o $OUT/prof_stem_$TAG $CMD > $OUT/ncu_stem_$TAG.log 2>&1
echo "stem capture rc=$?"
