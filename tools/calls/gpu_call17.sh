#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I spokennlp_b200/csrc -o /tmp/mma_shapes tools/micro/mma_shapes.cu 2>/dev/null
timeout 60 /tmp/mma_shapes | tee gpurun_out/r2q_mma_shapes.txt
timeout 300 python -m pytest tests/test_gpu_ponet.py -q -x -m gpu 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ponet_" --csv --log-file gpurun_out/r2q_ponet_times.csv python tools/prof_hbm.py > gpurun_out/r2q_ncu_ponet.log 2>&1; python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2q_ponet_times.csv')) if len(r)>5 and 'ponet' in ' '.join(r)]
for r in rows[:9]: print(r[4][:36], r[-1])
PY
