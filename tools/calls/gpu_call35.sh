#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r3i_gpu_suite.log 2>&1; tail -3 gpurun_out/r3i_gpu_suite.log
for v in pdl nopdl pdl nopdl; do
if [ $v = nopdl ]; then
timeout 300 python -c "
from spokennlp_b200 import lib
import os
lib.LIB_PATH = os.path.abspath('tools/micro/libb200enc_nopdl.so'); lib.is_stale = lambda: False
import runpy, sys
sys.argv = ['bench.py', '--no-padded', '--no-cpu-baseline', '--sustained-s', '0']
runpy.run_path('bench.py', run_name='__main__')" > gpurun_out/r3i_bench_$v.json 2> gpurun_out/r3i_bench_$v.err
else
timeout 300 python bench.py --no-padded --no-cpu-baseline --sustained-s 0 > gpurun_out/r3i_bench_$v.json 2> gpurun_out/r3i_bench_$v.err
fi
python - $v <<'PY'
import json, sys
v = sys.argv[1]
try:
    r=json.loads(open(f'gpurun_out/r3i_bench_{v}.json').read().strip().splitlines()[-1])
    print(v, round(r['value'],1), round(r['ms_per_step'],3), 'e2e', round(r['e2e']['value'],1), r['clocks']['sm_mhz'], r['final_loss'])
except Exception as e:
    print(v, 'no result', e); print(open(f'gpurun_out/r3i_bench_{v}.err').read()[-800:])
PY
done
