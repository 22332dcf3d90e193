#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fused.py tests/test_gpu_dropout.py tests/test_gpu_model.py tests/test_gpu_packing.py -q -x -m gpu 2>&1 | tail -4
timeout 120 python tools/gemm_stages.py 2>> gpurun_out/r3m_gemm_stages.err | tee gpurun_out/r3m_gemm_stages.jsonl | cut -c1-420
timeout 300 python tools/gemm_big.py 2> gpurun_out/r3m_gemm_big.err | tee gpurun_out/r3m_gemm_big.jsonl | cut -c1-200
timeout 200 python tools/kernel_timings.py 2> gpurun_out/r3m_kernel_timings.err | tee gpurun_out/r3m_kernel_timings.jsonl | grep "out_proj\|ffn_down  " | cut -c1-300
timeout 400 python bench.py --no-padded > gpurun_out/r3m_bench.json 2> gpurun_out/r3m_bench.err; python - <<'PY'
import json
r=json.loads(open('gpurun_out/r3m_bench.json').read().strip().splitlines()[-1])
print(r['value'], r['ms_per_step'], 'e2e', r['e2e']['value'], r['clocks'], r['encoder_flop_util']['frac_of_sustained'], r['roofline']['frac'])
print({k:round(v['ms']*1e3,1) for k,v in r['kernels'].items()})
print('sustained', r['sustained']['value'], r['sustained']['frac_of_sustained_peak'])
PY
