#!/bin/bash
# final-state evidence: launch list of one step, --set full of the roofline kernel and of the attention kernels with dropout
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 950 -c 240 --csv \
    --log-file gpurun_out/r3h_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --sustained-s 0 --no-padded > gpurun_out/r3h_ncu_bench.log 2>&1
tail -c 200 gpurun_out/r3h_ncu_bench.log
timeout 200 ncu --set full --clock-control none -k regex:gemm2_f16_kernel -c 4 -f -o gpurun_out/r3h_gelu python tools/roofline_traffic.py run > gpurun_out/r3h_ncu_gelu.log 2>&1; tail -2 gpurun_out/r3h_ncu_gelu.log
B200_ATTN_DROP=0.1 timeout 200 ncu --set full --clock-control none -k regex:"attn_fwd3|attn_bwd3" -c 2 -f -o gpurun_out/r3h_attn_drop python tools/prof_attn.py > gpurun_out/r3h_ncu_attn.log 2>&1; tail -2 gpurun_out/r3h_ncu_attn.log
