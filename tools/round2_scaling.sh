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
