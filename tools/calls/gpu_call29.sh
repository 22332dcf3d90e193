#!/bin/bash
mkdir -p gpurun_out
for l in "" tools/micro/libb200enc_nomma.so tools/micro/libb200enc_notma.so; do
B200_LIB=$l timeout 120 python tools/gemm_stages.py 2>> gpurun_out/r3c_gemm_diag.err | tee -a gpurun_out/r3c_gemm_diag.jsonl
done
