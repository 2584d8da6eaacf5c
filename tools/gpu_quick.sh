#!/bin/bash
# quick GPU check of a kernel change: the C2 / fused parity tests, then one bench line (no CPU leg)
mkdir -p gpurun_out
tag=${1:-q}
python -m pytest tests -m gpu -x -q -k "c2 or C2 or fused or smoke or solve_host" > gpurun_out/r2_${tag}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_${tag}_tests.log
python bench.py --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r2_${tag}_bench.json 2> gpurun_out/r2_${tag}_bench.err
