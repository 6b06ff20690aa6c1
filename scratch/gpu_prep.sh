#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -m gpu -q --timeout 150 > gpurun_out/prep_tests.log 2>&1; echo "tests rc=$?"; grep -n "^E   \|FAILED\|passed\|failed" gpurun_out/prep_tests.log | head
timeout 300 python - <<'PY'
import sys, copy, torch
sys.path.insert(0, '.')
import bench
sys.argv = sys.argv[:1]
args = bench.parse(); dev = torch.device("cuda:0")
scene = bench.host_scene(args.shape, args.n_class)
for _ in range(2): print(bench._native_iteration(args, dev, scene, 2000))
PY
timeout 200 python scratch/cfg1_time.py 2>&1 | cut -c1-60
