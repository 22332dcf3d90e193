#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_packing.py -q > gpurun_out/r2h_packing.log 2>&1; tail -8 gpurun_out/r2h_packing.log; grep -n "^E " gpurun_out/r2h_packing.log | head -10
timeout 300 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2h_bench.json").read().strip().splitlines()[-1])
    print(round(d["value"], 1), "seq/s", round(d["ms_per_step"], 3), "ms; sustained", d["sustained"]["value"], d["sustained"]["frac_of_sustained_peak"])
    print("padded", json.dumps(d["padded"]))
except Exception as e:
    print("no result:", e); print(open("gpurun_out/r2h_bench.err").read()[-2000:])
PY
