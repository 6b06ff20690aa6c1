#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scratch/native_iter.py 6 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/native_launches.csv python scratch/native_iter.py 6 > gpurun_out/native_ncu.log 2>&1; echo rc=$?
