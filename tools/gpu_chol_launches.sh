#!/bin/bash
# launch lists (ncu, serialised) of one Cholesky factorisation of the tiled path: C4 at batch 1000, C5
mkdir -p gpurun_out
tag=${1:-chol}
for cfg in C4 C5; do
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_${tag}_launches_${cfg}.csv python tools/bench_steps.py $cfg --reps 1 --only 8 > gpurun_out/r2_${tag}_ncu_${cfg}.log 2>&1
done
