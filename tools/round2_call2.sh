#!/bin/bash
# GPU call 2 of round 2: the sixteen-softmax-warp attention kernels (attn_fwd4 / attn_bwd4) — parity, kernel A/B, bench A/B.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_experimental.py -q -x -k "warp_elected" > gpurun_out/r2b_attn_parity.log 2>&1; tail -5 gpurun_out/r2b_attn_parity.log
timeout 400 python -m pytest tests/test_gpu_experimental.py -q -k "training_step and (bwd16 or fwd16)" > gpurun_out/r2b_step_parity.log 2>&1; tail -5 gpurun_out/r2b_step_parity.log
timeout 200 python tools/variants_ab.py > gpurun_out/r2b_variants_ab.jsonl 2> gpurun_out/r2b_variants_ab.err; tail -2 gpurun_out/r2b_variants_ab.jsonl
for exp in "default" "resadd,delta,elect,bwd16" "resadd,delta,elect,fwd16" "resadd,delta,elect,bwd16,fwd16"; do
  tag=${exp//,/_}
  B200_EXP="$exp" timeout 150 python bench.py --no-cpu-baseline --steps 12 > gpurun_out/r2b_bench_$tag.json 2> gpurun_out/r2b_bench_$tag.err
  python - "$tag" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2b_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"], 1), "seq/s", round(d["ms_per_step"], 3), "ms  e2e", round(d["e2e"]["value"], 1), "loss", d["final_loss"], d["config"]["opt_in_variants"])
except Exception as e:
    print(sys.argv[1], "no result:", e)
PY
done
timeout 420 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_experimental.py > gpurun_out/r2b_gpu_suite.log 2>&1; tail -4 gpurun_out/r2b_gpu_suite.log
