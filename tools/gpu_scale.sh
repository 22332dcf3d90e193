#!/bin/bash
# Data-parallel scaling runs on N GPUs of one box:  gpurun --gpus N --timeout 900 -- 'bash tools/gpu_scale.sh N [variant ...]'
# Each variant is "NAME:ENV=VAL,ENV=VAL"; one bench.py run (torchrun, one rank per GPU) per variant, bounded by its own timeout
# and bench.py's watchdog, so a stuck collective costs one variant, not the call.
N=${1:-2}; shift
mkdir -p gpurun_out
VARIANTS=("$@")
[ ${#VARIANTS[@]} -eq 0 ] && VARIANTS=("graph:" "graph_ctas4:B200_COMM_CTAS=4")
for v in "${VARIANTS[@]}"; do
  name=${v%%:*}; envs=${v#*:}
  ( IFS=','; for kv in $envs; do [ -n "$kv" ] && export "$kv"; done
    B200_BENCH_WATCHDOG_S=200 timeout 260 python bench.py --gpus $N --steps 30 --no-cpu-baseline --sustained-s 0 --no-padded \
      > gpurun_out/scale_n${N}_${name}.json 2> gpurun_out/scale_n${N}_${name}.err )
  python - "$N" "$name" <<'PY'
import json, sys
n, name = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(f"gpurun_out/scale_n{n}_{name}.json").read().strip().splitlines()[-1])
    print(f"N={n} {name}: {d['value']:.1f} seq/s  {d['ms_per_step']:.3f} ms/step  per-GPU {d['value']/int(n):.1f}  e2e {d['e2e']['value']:.1f}  graph={d['config']['cuda_graph']}  nccl={json.dumps(d.get('nccl'))[:400]}")
except Exception as e:
    print(f"N={n} {name}: no result ({e})"); print(open(f"gpurun_out/scale_n{n}_{name}.err").read()[-1200:])
PY
done
