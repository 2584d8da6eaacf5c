#!/bin/bash
# bench lines (device-resident value only matters) under a list of environment settings: gpu_env_sweep.sh tag "A=1 B=2" "A=0" ...
tag=$1; shift
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  env $envs python bench.py --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r2_${tag}_$i.json 2> gpurun_out/r2_${tag}_$i.err
  echo "$envs :: $(grep -o '"value": [0-9.]*' gpurun_out/r2_${tag}_$i.json | head -1) e2e $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r2_${tag}_$i.json)" >> gpurun_out/r2_${tag}_summary.txt
  i=$((i+1))
done
