#!/bin/bash
# GPU call 4: packed-fp32 GELU / dropout epilogues, cls head with contiguous row ranges; launch list of one step.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2d_gpu_suite.log 2>&1; tail -5 gpurun_out/r2d_gpu_suite.log
timeout 200 python tools/kernel_timings.py > gpurun_out/r2d_kernel_timings.jsonl 2> gpurun_out/r2d_kernel_timings.err; cat gpurun_out/r2d_kernel_timings.jsonl; tail -3 gpurun_out/r2d_kernel_timings.err
timeout 150 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2d_bench.json").read().strip().splitlines()[-1])
    print(round(d["value"], 1), "seq/s", round(d["ms_per_step"], 3), "ms  e2e", round(d["e2e"]["value"], 1), "loss", d["final_loss"], d["roofline"]["frac"], d["clocks"])
except Exception as e:
    print("no result:", e); print(open("gpurun_out/r2d_bench.err").read()[-1500:])
PY
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 330 --csv \
    --log-file gpurun_out/r2d_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2d_ncu_bench.log 2>&1
tail -c 300 gpurun_out/r2d_ncu_bench.log
