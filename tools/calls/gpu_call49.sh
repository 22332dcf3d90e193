#!/bin/bash
mkdir -p gpurun_out
for l in "" tools/micro/libb200enc_lnb192.so "" tools/micro/libb200enc_lnb192.so; do
B200_LIB=$l timeout 120 python tools/ln_timing.py 2>> gpurun_out/r3q_ln.err | tee -a gpurun_out/r3q_ln.jsonl
done
