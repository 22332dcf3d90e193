#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none -k regex:"ln_|embed_ln|cls_head|ponet_|ce_stats" -f -o gpurun_out/r3t_hbm python tools/prof_hbm.py > gpurun_out/r3t_ncu_hbm.log 2>&1; tail -2 gpurun_out/r3t_ncu_hbm.log
