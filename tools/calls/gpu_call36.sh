#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r3u_gpu_suite.log 2>&1; tail -3 gpurun_out/r3u_gpu_suite.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 500 python bench.py > gpurun_out/r3u_bench.json 2> gpurun_out/r3u_bench.err; python - <<'PY'
import json
r=json.loads(open('gpurun_out/r3u_bench.json').read().strip().splitlines()[-1])
print(r['value'], r['ms_per_step'], 'e2e', r['e2e']['value'], r['clocks'], r['encoder_flop_util']['frac_of_sustained'], r['roofline']['frac'], r['roofline']['traffic'])
print('sustained', r['sustained']['value'], r['sustained']['frac_of_sustained_peak'], 'padded', r['padded']['value'], r['padded']['frac_of_ideal'], 'cpu', r['cpu_baseline'])
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
