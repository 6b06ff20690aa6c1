#!/bin/bash
mkdir -p gpurun_out
DNS_SLAM_B200_LIB=$PWD/dns_slam_b200/libdns_slam_b200_ablate.so timeout 300 python scratch/phase_clk_fm.py > gpurun_out/phase_clk_fm.log 2>&1; echo "rc=$?"; tail -18 gpurun_out/phase_clk_fm.log
timeout 400 python -m pytest tests/test_gpu_featmerge.py tests/test_gpu_framestep.py tests/test_gpu_loops.py tests/test_gpu_parity_configs.py tests/test_gpu_inference.py -q --timeout 150 > gpurun_out/fm_tests.log 2>&1; echo "tests rc=$?"; grep -n "^E   \|FAILED\|passed\|failed" gpurun_out/fm_tests.log | head -20
timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu > gpurun_out/fm_bench.json 2> gpurun_out/fm_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/fm_bench.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/fm_bench.json'))
print(b['value'], b['ms_per_step'], b['e2e']['value'], {k: round(v,3) for k,v in b['roofline']['phase_ms_per_step'].items()})
PY
