#!/bin/bash
# GPU call 3 of round 2: trimmed library (elect-only attention with packed fp32 math + quad dropout masks, fused dGELU column sum,
# new cls head): full GPU suite, kernel timings, bench A/B of the schedule switches, attention ncu capture.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2c_gpu_suite.log 2>&1; tail -12 gpurun_out/r2c_gpu_suite.log
timeout 200 python tools/kernel_timings.py > gpurun_out/r2c_kernel_timings.jsonl 2> gpurun_out/r2c_kernel_timings.err; cat gpurun_out/r2c_kernel_timings.jsonl; tail -3 gpurun_out/r2c_kernel_timings.err
for exp in "default" "resadd,delta"; do
  tag=${exp//,/_}
  B200_EXP="$exp" timeout 150 python bench.py --no-cpu-baseline --steps 12 > gpurun_out/r2c_bench_$tag.json 2> gpurun_out/r2c_bench_$tag.err
  python - "$tag" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2c_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"], 1), "seq/s", round(d["ms_per_step"], 3), "ms  e2e", round(d["e2e"]["value"], 1), "loss", d["final_loss"], d["config"].get("disabled_default_variants"))
except Exception as e:
    print(sys.argv[1], "no result:", e)
    print(open(f"gpurun_out/r2c_bench_{sys.argv[1]}.err").read()[-1500:])
PY
done
B200_ATTN_DROP=0.1 timeout 280 ncu --set full --clock-control none --import-source on -k regex:"attn_(fwd3|bwd3)_kernel" -c 2 \
  -f -o gpurun_out/r2c_attn_drop python tools/prof_attn.py > gpurun_out/r2c_ncu_attn_drop.log 2>&1; tail -2 gpurun_out/r2c_ncu_attn_drop.log
