#!/bin/bash
# round 2, GPU call a: new kernels first (bounded), then the whole GPU suite, then a short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r02a_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_featmerge.py tests/test_gpu_framestep.py -x -q > gpurun_out/r02a_new.log 2>&1
echo "new rc=$?" | tee -a gpurun_out/r02a_new.log
tail -30 gpurun_out/r02a_new.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02a_all.log 2>&1
echo "all rc=$?" | tee -a gpurun_out/r02a_all.log
tail -40 gpurun_out/r02a_all.log
timeout 900 python bench.py --steps 5 --warmup 2 --no-extra --no-cpu > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo "bench rc=$?"
tail -5 gpurun_out/r02a_bench.err
cat gpurun_out/r02a_bench.json
