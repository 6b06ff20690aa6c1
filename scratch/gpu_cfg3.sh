#!/bin/bash
mkdir -p gpurun_out
timeout 400 python scratch/cfg3_prof.py scannet 26 > gpurun_out/cfg3_prof.log 2>&1; echo "rc=$?"; head -70 gpurun_out/cfg3_prof.log
