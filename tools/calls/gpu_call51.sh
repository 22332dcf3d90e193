#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_dropout.py tests/test_gpu_cross.py tests/test_gpu_ponet.py -q -x -m gpu 2>&1 | tail -3
timeout 120 python tools/ln_timing.py 2>> gpurun_out/r3s_ln.err | tee -a gpurun_out/r3s_ln.jsonl | cut -c1-220
timeout 400 python bench.py --no-padded > gpurun_out/r3s_bench.json 2> gpurun_out/r3s_bench.err; python - <<'PY'
import json
r=json.loads(open('gpurun_out/r3s_bench.json').read().strip().splitlines()[-1])
print(r['value'], r['ms_per_step'], 'e2e', r['e2e']['value'], r['clocks'], r['encoder_flop_util']['frac_of_sustained'], r['roofline']['frac'])
print({k:round(v['ms']*1e3,1) for k,v in r['kernels'].items()})
print('sustained', r['sustained']['value'], r['sustained']['frac_of_sustained_peak'])
PY
