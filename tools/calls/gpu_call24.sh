#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2x_gpu_suite.log 2>&1; tail -3 gpurun_out/r2x_gpu_suite.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; python - <<'PY'
import json
r=json.loads(open('gpurun_out/r2x_bench.json').read().strip().splitlines()[-1])
print(r['value'], r['ms_per_step'], 'e2e', r['e2e']['value'], 'roofline', r['roofline'], 'sust', r.get('sustained'), 'padded', r.get('padded'))
print({k:(round(v.get('us',0),1) if isinstance(v,dict) else v) for k,v in r.get('kernels',{}).items()})
PY
timeout 200 python tools/kernel_timings.py 2> gpurun_out/r2x_kernel_timings.err > gpurun_out/r2x_kernel_timings.jsonl; grep attention gpurun_out/r2x_kernel_timings.jsonl
