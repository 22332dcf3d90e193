#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_dropout.py tests/test_gpu_packing.py tests/test_gpu_cross.py tests/test_gpu_fused.py -q -x -m gpu 2>&1 | tail -3
timeout 120 python tools/attn_trace.py bwd 0.1 > gpurun_out/r2w_trace_bwd_0.1.txt 2> gpurun_out/r2w_trace_bwd.err; tail -16 gpurun_out/r2w_trace_bwd_0.1.txt
timeout 300 python tools/attn_scaling.py 2> gpurun_out/r2w_attn_scaling.err | tee gpurun_out/r2w_attn_scaling.jsonl | head -4
