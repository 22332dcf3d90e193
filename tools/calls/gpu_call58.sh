#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -m gpu -k "layernorm or embed" 2>&1 | tail -2
for i in 1 2; do timeout 120 python tools/ln_timing.py 2>> gpurun_out/r3v_ln.err | tee -a gpurun_out/r3v_ln.jsonl | cut -c1-200; done
