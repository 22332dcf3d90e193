#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I spokennlp_b200/csrc -o /tmp/softmax_body tools/micro/softmax_body.cu 2>/dev/null
timeout 120 /tmp/softmax_body | tee gpurun_out/r2l_softmax_body.txt
timeout 300 python tools/attn_scaling.py 2> gpurun_out/r2l_attn_scaling.err | tee gpurun_out/r2l_attn_scaling.jsonl
