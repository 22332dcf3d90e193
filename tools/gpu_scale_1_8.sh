#!/bin/bash
# N = 1 and N = 8 back to back on ONE 8-GPU box (default settings):  gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_scale_1_8.sh'
mkdir -p gpurun_out
for N in 1 8; do bash tools/gpu_scale.sh $N "default:"; done
