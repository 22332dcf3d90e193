#!/bin/bash
mkdir -p gpurun_out
timeout 500 python bench.py > gpurun_out/r3l_bench.json 2> gpurun_out/r3l_bench.err; python - <<'PY'
import json
r=json.loads(open('gpurun_out/r3l_bench.json').read().strip().splitlines()[-1])
print(r['value'], r['ms_per_step'], 'e2e', r['e2e']['value'], r['clocks'], r['encoder_flop_util']['frac_of_sustained'], 'roofline', r['roofline']['achieved'], r['roofline']['frac'])
print({k:round(v['ms']*1e3,1) for k,v in r['kernels'].items()})
print('sustained', r['sustained']['value'], r['sustained']['frac_of_sustained_peak'], 'padded', r['padded']['value'], r['padded']['frac_of_ideal'])
PY
tail -3 gpurun_out/r3l_bench.err
