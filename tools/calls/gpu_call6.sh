#!/bin/bash
# GPU call 6: new bench line (sustained + padded records, kernels as the step runs them), GELU-GEMM ncu capture for roofline.traffic,
# ncu of the HBM-bound kernels (cls head, embeddings, LayerNorm).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "cls_head" > gpurun_out/r2f_cls.log 2>&1; tail -3 gpurun_out/r2f_cls.log
timeout 300 python bench.py --steps 20 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 3000 gpurun_out/r2f_bench.json; tail -5 gpurun_out/r2f_bench.err
timeout 200 ncu --set full --clock-control none -k regex:gemm2_f16_kernel -c 4 -f -o gpurun_out/r2f_gelu python tools/roofline_traffic.py run > gpurun_out/r2f_ncu_gelu.log 2>&1; tail -2 gpurun_out/r2f_ncu_gelu.log
timeout 200 ncu --set full --clock-control none -k regex:"cls_head|embed_ln|ln_fwd|ln_bwd2" -c 12 -f -o gpurun_out/r2f_hbm python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --sustained-s 0 --no-padded > gpurun_out/r2f_ncu_hbm.log 2>&1; tail -c 300 gpurun_out/r2f_ncu_hbm.log
