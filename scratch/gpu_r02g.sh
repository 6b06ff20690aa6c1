#!/bin/bash
mkdir -p gpurun_out
( time python bench.py > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err ) 2>&1 | tail -3
echo "bench rc=$?"; tail -3 gpurun_out/r02g_bench.err
( time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02g_ref.json 2> gpurun_out/r02g_ref.err ) 2>&1 | tail -3
cat gpurun_out/r02g_ref.json | cut -c1-400
bash profiles/run_ncu.sh r02g
