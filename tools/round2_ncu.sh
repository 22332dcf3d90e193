#!/bin/bash
# ncu captures for round 2 (one B200; ncu replays each kernel ~40x, keep the launch counts small):
#   gpurun --timeout 900 -- 'bash tools/round2_ncu.sh'
# 1. attention forward / backward, default and hand-off variants, --set full with source + stall sampling
#    -> read here with tools/ncu_hot.py / tools/ncu_lines.py
# 2. the GEMM family at the bench shapes incl. the opt-in epilogues (tools/variants_ab.py launches them)
# 3. launch list of one bench step (shares per kernel) for the default path and for the best variant set ($BEST_EXP)
mkdir -p gpurun_out
for v in 0 1 3; do
  B200_ATTN_VARIANT=$v timeout 280 ncu --set full --clock-control none --import-source on -k regex:"attn_(fwd3|bwd3)_kernel" -c 2 \
    -f -o gpurun_out/r2_attn_v$v python tools/prof_attn.py > gpurun_out/r2_ncu_attn_v$v.log 2>&1; tail -2 gpurun_out/r2_ncu_attn_v$v.log
done
timeout 280 ncu --set full --clock-control none --import-source on -k regex:"gemm2_f16_kernel" -s 8 -c 12 \
  -f -o gpurun_out/r2_gemm_variants python tools/variants_ab.py > gpurun_out/r2_ncu_gemm.log 2>&1; tail -2 gpurun_out/r2_ncu_gemm.log
for exp in "" "${BEST_EXP:-resadd,delta}"; do
  tag=${exp//,/_}; tag=${tag:-default}
  B200_EXP="$exp" timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 330 --csv \
    --log-file gpurun_out/r2_launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2_ncu_bench_$tag.log 2>&1
  tail -c 200 gpurun_out/r2_ncu_bench_$tag.log
done
