#!/bin/bash
# N = 1, 2, 4 back to back on ONE 4-GPU box (default settings):  gpurun --gpus 4 --timeout 900 -- 'bash tools/gpu_scale_1_2_4.sh'
mkdir -p gpurun_out
for N in 1 2 4; do bash tools/gpu_scale.sh $N "default:"; done
