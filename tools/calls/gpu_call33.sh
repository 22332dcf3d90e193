#!/bin/bash
mkdir -p gpurun_out
timeout 500 python tools/config_sweeps.py > gpurun_out/r3g_config_sweeps.log 2>&1; tail -12 gpurun_out/r3g_config_sweeps.log | cut -c1-300
cp gpurun_out/config_sweeps.jsonl gpurun_out/config_sweeps_r02.jsonl 2>/dev/null
