#!/bin/bash
mkdir -p gpurun_out
for w in fwd bwd; do for p in 0.1; do
timeout 120 python tools/attn_trace.py $w $p > gpurun_out/r2m_trace_${w}_${p}.txt 2> gpurun_out/r2m_trace_${w}.err; tail -22 gpurun_out/r2m_trace_${w}_${p}.txt; tail -3 gpurun_out/r2m_trace_${w}.err
done; done
