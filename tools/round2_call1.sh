#!/bin/bash
# First GPU call of round 2 (one B200, ~12 min of box time):
#   gpurun --timeout 900 -- 'bash tools/round2_call1.sh'
# 1. the default GPU suite, 2. the opt-in variants' tests, each node in its own process (a trapped kernel cannot hide the
# others), 3. kernel-level A/B, 4. bench.py with each variant set on the same box.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_suite.log 2>&1; tail -15 gpurun_out/r2_gpu_suite.log
# one process first (fast); only if something fails or traps, every node again in its own process
if B200_RUN_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_gpu_experimental.py -q > gpurun_out/r2_experimental.log 2>&1; then
  tail -2 gpurun_out/r2_experimental.log
else
  tail -5 gpurun_out/r2_experimental.log
  B200_RUN_EXPERIMENTAL=1 timeout 600 python tools/gpu_isolated.py tests/test_gpu_experimental.py --timeout 60 > gpurun_out/r2_experimental_isolated.log 2>&1
  grep -E "^(PASS|FAIL)|passed" gpurun_out/r2_experimental_isolated.log | tail -40
fi
# memcheck over the new template instantiations (small shapes; SURVEY §5: sanitizer pass per kernel)
B200_RUN_EXPERIMENTAL=1 timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_experimental.py -q -x \
  -k "256-256-64 or 1024-768-768 or 300-768-1536 or 2-128-2 or 3-300-4" > gpurun_out/r2_sanitizer_memcheck.log 2>&1; tail -4 gpurun_out/r2_sanitizer_memcheck.log
timeout 120 python tools/variants_ab.py > gpurun_out/r2_variants_ab.jsonl 2> gpurun_out/r2_variants_ab.err; cat gpurun_out/r2_variants_ab.jsonl
for exp in "" "resadd" "delta" "elect" "ewait" "resadd,delta,ewait" "streamk,delta,ewait"; do
  tag=${exp//,/_}; tag=${tag:-default}
  B200_EXP="$exp" timeout 150 python bench.py --no-cpu-baseline --steps 12 > gpurun_out/r2_bench_$tag.json 2> gpurun_out/r2_bench_$tag.err
  python - "$tag" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"], 1), "seq/s", round(d["ms_per_step"], 3), "ms  e2e", round(d["e2e"]["value"], 1), "loss", d["final_loss"], d["config"]["opt_in_variants"])
except Exception as e:
    print(sys.argv[1], "no result:", e)
PY
done
# the persistent attention kernels as they run in training (dropout 0.1 on the probabilities): --set full + source-level stall sampling
B200_ATTN_DROP=0.1 timeout 280 ncu --set full --clock-control none --import-source on -k regex:"attn_(fwd3|bwd3)_kernel" -c 2 \
  -f -o gpurun_out/r2_attn_drop_v0 python tools/prof_attn.py > gpurun_out/r2_ncu_attn_drop_v0.log 2>&1; tail -2 gpurun_out/r2_ncu_attn_drop_v0.log
