#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_dropout.py tests/test_gpu_packing.py tests/test_gpu_cross.py tests/test_gpu_fused.py -q -x -m gpu 2>&1 | tail -6
for w in fwd bwd; do
timeout 120 python tools/attn_trace.py $w 0.1 > gpurun_out/r2t_trace_${w}_0.1.txt 2> gpurun_out/r2t_trace_${w}.err; tail -14 gpurun_out/r2t_trace_${w}_0.1.txt
done
timeout 300 python tools/attn_scaling.py 2> gpurun_out/r2t_attn_scaling.err | tee gpurun_out/r2t_attn_scaling.jsonl | head -4
