#!/bin/bash
# The driver's scaling sequence on ONE 8-GPU box: N = 1, 2, 4, 8 back to back (default settings), then the A/B at N = 8 without
# the capped background communicator.   gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_scale_all.sh'
mkdir -p gpurun_out
for N in 1 2 4 8; do bash tools/gpu_scale.sh $N "default:"; done
bash tools/gpu_scale.sh 8 "ctas0:B200_COMM_CTAS=0" "eager:B200_DP_GRAPH=0"
