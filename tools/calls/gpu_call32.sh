#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fused.py tests/test_gpu_dropout.py tests/test_gpu_model.py -q -x -m gpu 2>&1 | tail -3
timeout 120 python tools/gemm_stages.py 2>> gpurun_out/r3f_gemm_stages.err | tee gpurun_out/r3f_gemm_stages.jsonl | cut -c1-420
timeout 200 python tools/kernel_timings.py 2> gpurun_out/r3f_kernel_timings.err | tee gpurun_out/r3f_kernel_timings.jsonl | grep "out_proj  \|ffn_down  "
timeout 400 python bench.py --no-padded > gpurun_out/r3f_bench.json 2> gpurun_out/r3f_bench.err; python - <<'PY'
import json
r=json.loads(open('gpurun_out/r3f_bench.json').read().strip().splitlines()[-1])
print(r['value'], r['ms_per_step'], 'e2e', r['e2e']['value'], r['clocks'], r['encoder_flop_util']['frac_of_sustained'], r['roofline']['frac'])
print({k:round(v['ms']*1e3,1) for k,v in r['kernels'].items()})
PY
