#!/bin/bash
# Scaling check of the default data-parallel path (round 1 measured 1 and 2 GPUs only):
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/round2_scaling.sh'
# Every run has its own timeout and the bench's watchdog is set BELOW it, so a stuck collective ends with a JSON line that
# says so instead of eating the call.  NCCL warnings go to the .err files.
mkdir -p gpurun_out
for n in 2 4 8; do
  NCCL_DEBUG=WARN B200_BENCH_WATCHDOG_S=170 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
    --master-port $((29600 + n)) bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r2_scale_n$n.json 2> gpurun_out/r2_scale_n$n.err
  echo "n=$n rc=$?"; tail -c 400 gpurun_out/r2_scale_n$n.json | cut -c1-400; grep -i "warn\|error\|watchdog" gpurun_out/r2_scale_n$n.err | head -5
done
# Second stage (only with STAGE2=1): the captured step and the CTA-capped background communicator beyond 2 GPUs — both were
# validated on 2 GPUs only in round 1.
if [ "${STAGE2:-0}" = "1" ]; then
  for cfg in "graph 1 0" "graph_comm4 1 4" "eager_comm4 0 4"; do
    set -- $cfg
    for n in 4 8; do
      NCCL_DEBUG=WARN B200_DP_GRAPH=$2 B200_COMM_CTAS=$3 B200_BENCH_WATCHDOG_S=170 timeout 200 python -m torch.distributed.run --nnodes=1 \
        --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n + 10 * $3 + 100 * $2)) bench.py --gpus $n --steps 20 --warmup 3 \
        > gpurun_out/r2_scale_$1_n$n.json 2> gpurun_out/r2_scale_$1_n$n.err
      echo "$1 n=$n rc=$?"; tail -c 300 gpurun_out/r2_scale_$1_n$n.json | cut -c1-300
    done
  done
  # NCCL's own CTA cap on the single default communicator (no second communicator, no SM reservation)
  for c in 4 8 16; do
    NCCL_DEBUG=WARN NCCL_MAX_CTAS=$c B200_BENCH_WATCHDOG_S=170 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 \
      --master-addr 127.0.0.1 --master-port $((29900 + c)) bench.py --gpus 8 --steps 20 --warmup 3 \
      > gpurun_out/r2_scale_maxctas${c}_n8.json 2> gpurun_out/r2_scale_maxctas${c}_n8.err
    echo "NCCL_MAX_CTAS=$c n=8 rc=$?"; tail -c 300 gpurun_out/r2_scale_maxctas${c}_n8.json | cut -c1-300
  done
fi
