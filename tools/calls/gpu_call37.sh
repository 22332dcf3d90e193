#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_ponet.py -q -m gpu 2>&1 | tail -8
