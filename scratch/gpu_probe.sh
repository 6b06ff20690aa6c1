#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scratch/e2e_probe.py > gpurun_out/e2e_probe.log 2>&1; echo "rc=$?"; head -60 gpurun_out/e2e_probe.log
