#!/bin/bash
# GPU call 7: packed rows (varlen attention + packed training step), heads after the compaction refactor, full suite.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_packing.py tests/test_gpu_heads.py -q > gpurun_out/r2g_packing.log 2>&1; tail -25 gpurun_out/r2g_packing.log
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_packing.py --deselect tests/test_gpu_heads.py > gpurun_out/r2g_gpu_suite.log 2>&1; tail -5 gpurun_out/r2g_gpu_suite.log
