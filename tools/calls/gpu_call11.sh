#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_ponet.py tests/test_gpu_fused.py tests/test_mmvts_encoders.py tests/test_gpu_cross.py -q -s -m gpu > gpurun_out/r2k_tests.log 2>&1; tail -5 gpurun_out/r2k_tests.log; grep -n "bf16 arm\|^E " gpurun_out/r2k_tests.log | head
timeout 200 python tools/kernel_timings.py 2> gpurun_out/r2k_kernel_timings.err | tee gpurun_out/r2k_kernel_timings.jsonl | tail -3
timeout 300 ncu --set full --clock-control none -k regex:"ponet_" -f -o gpurun_out/r2k_ponet python tools/prof_hbm.py > gpurun_out/r2k_ncu_ponet.log 2>&1; tail -2 gpurun_out/r2k_ncu_ponet.log
