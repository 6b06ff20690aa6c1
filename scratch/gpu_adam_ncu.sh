#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_adam_multi -s 30 -c 1 -o gpurun_out/adam python scratch/cfg1_time.py > gpurun_out/adam_ncu.log 2>&1; echo rc=$?
