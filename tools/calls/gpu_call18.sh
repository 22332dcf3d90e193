#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I spokennlp_b200/csrc -o /tmp/mma_shapes tools/micro/mma_shapes.cu 2>/dev/null
timeout 60 /tmp/mma_shapes | tee gpurun_out/r2r_mma_shapes.txt
timeout 120 python tools/attn_trace.py bwd 0.1 > gpurun_out/r2r_trace_bwd_0.1.txt 2> gpurun_out/r2r_trace_bwd.err; grep "^warp 13" gpurun_out/r2r_trace_bwd_0.1.txt | cut -c1-600
