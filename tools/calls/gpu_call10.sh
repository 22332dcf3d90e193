#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2j_gpu_suite.log 2>&1; tail -5 gpurun_out/r2j_gpu_suite.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none -k regex:"ln_|embed_ln|cls_head|ponet_|ce_stats" -f -o gpurun_out/r2j_hbm python tools/prof_hbm.py > gpurun_out/r2j_ncu_hbm.log 2>&1; tail -2 gpurun_out/r2j_ncu_hbm.log
