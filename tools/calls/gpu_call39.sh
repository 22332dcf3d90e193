#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_graphs.py tests/test_gpu_cross.py tests/test_mmvts_encoders.py -q -x -m gpu 2>&1 | tail -15
timeout 500 python tools/config_sweeps.py > gpurun_out/r3k_config_sweeps.log 2>&1; grep '"config": 4' gpurun_out/r3k_config_sweeps.log | cut -c1-260; tail -3 gpurun_out/r3k_config_sweeps.log | cut -c1-300
cp gpurun_out/config_sweeps.jsonl gpurun_out/config_sweeps_r02.jsonl 2>/dev/null
