#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_dropout.py tests/test_gpu_packing.py tests/test_gpu_cross.py -q -x -m gpu 2>&1 | tail -3
timeout 300 python tools/attn_scaling.py 2> gpurun_out/r2y_attn_scaling.err | tee gpurun_out/r2y_attn_scaling.jsonl | head -4 | cut -c1-200
for l in "" tools/micro/libb200enc_s3.so tools/micro/libb200enc_s4.so tools/micro/libb200enc_s6.so; do
B200_LIB=$l timeout 120 python tools/gemm_stages.py 2>> gpurun_out/r2y_gemm_stages.err | tee -a gpurun_out/r2y_gemm_stages.jsonl
done
