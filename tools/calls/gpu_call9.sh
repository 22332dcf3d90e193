#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fused.py -q -k "dq_half or dq16" > gpurun_out/r2i_dq16.log 2>&1; tail -6 gpurun_out/r2i_dq16.log; grep -n "^E " gpurun_out/r2i_dq16.log | head
for exp in "default" "resadd,delta,colsum,dq16"; do
  tag=${exp//,/_}
  B200_EXP="$exp" timeout 200 python bench.py --no-cpu-baseline --steps 20 --no-padded > gpurun_out/r2i_bench_$tag.json 2> gpurun_out/r2i_bench_$tag.err
  python - "$tag" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2i_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"], 1), "seq/s", round(d["ms_per_step"], 3), "ms  sustained", round(d["sustained"]["value"],1), round(d["sustained"]["frac_of_sustained_peak"],4), "loss", d["final_loss"], d["config"]["opt_in_variants"])
except Exception as e:
    print(sys.argv[1], "no result:", e); print(open(f"gpurun_out/r2i_bench_{sys.argv[1]}.err").read()[-1500:])
PY
done
