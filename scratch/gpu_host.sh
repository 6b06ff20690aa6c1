#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scratch/host_prof.py > gpurun_out/host_prof.log 2>&1; echo "rc=$?"; head -50 gpurun_out/host_prof.log
