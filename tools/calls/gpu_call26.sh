#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gemm_big.py 2> gpurun_out/r2z_gemm_big.err | tee gpurun_out/r2z_gemm_big.jsonl
