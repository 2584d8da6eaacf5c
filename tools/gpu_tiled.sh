#!/bin/bash
# tiled path: C4 / C5 tests, per-kernel step timings and bench lines with and without the TMA feed
tag=${1:-t}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "c4 or c5 or tma" > gpurun_out/r2_${tag}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_${tag}_tests.log
for cfg in C4 C5; do
  python tools/bench_steps.py $cfg --reps 3 --only 7,8 > gpurun_out/r2_${tag}_steps_${cfg}_tma.txt 2>&1
  SOCP_B200_NO_TMA=1 python tools/bench_steps.py $cfg --reps 3 --only 7,8 > gpurun_out/r2_${tag}_steps_${cfg}_cpasync.txt 2>&1
  python bench.py --config $cfg --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r2_${tag}_bench_${cfg}.json 2> gpurun_out/r2_${tag}_bench_${cfg}.err
done
