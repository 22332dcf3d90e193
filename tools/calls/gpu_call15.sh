#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_dropout.py tests/test_gpu_packing.py tests/test_gpu_cross.py -q -x -m gpu 2>&1 | tail -4
timeout 120 python tools/attn_trace.py fwd 0.1 > gpurun_out/r2o_trace_fwd_0.1.txt 2> gpurun_out/r2o_trace_fwd.err; tail -14 gpurun_out/r2o_trace_fwd_0.1.txt
timeout 300 python tools/attn_scaling.py 2> gpurun_out/r2o_attn_scaling.err | tee gpurun_out/r2o_attn_scaling.jsonl | head -4
