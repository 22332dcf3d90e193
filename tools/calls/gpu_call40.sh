#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_graphs.py tests/test_gpu_model.py tests/test_gpu_ponet.py tests/test_gpu_cross.py tests/test_wrapper_heads.py tests/test_gpu_dropout.py -q -x -m gpu 2>&1 | tail -15
