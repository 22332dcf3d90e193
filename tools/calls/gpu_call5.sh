#!/bin/bash
# GPU call 5: loss-head kernels (f1) against the reference-minted goldens; cls head v3; full suite.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_heads.py -q > gpurun_out/r2e_heads.log 2>&1; tail -25 gpurun_out/r2e_heads.log
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_heads.py > gpurun_out/r2e_gpu_suite.log 2>&1; tail -5 gpurun_out/r2e_gpu_suite.log
timeout 200 python tools/kernel_timings.py 2> gpurun_out/r2e_kernel_timings.err | tail -2
