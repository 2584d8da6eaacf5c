#!/bin/bash
# weak-scaling lines on N GPUs of one box: gpu_scale.sh N   (run under `gpurun --gpus N`)
N=${1:-2}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
run --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_final_bench_c2_${N}gpu.json 2> gpurun_out/r02_final_bench_c2_${N}gpu.err
if [ "$N" = "8" ]; then
  run --steps 5 --warmup 3 --no-cpu-baseline --batch 12500 > gpurun_out/r02_final_bench_c2_8gpu_100k.json 2> gpurun_out/r02_final_bench_c2_8gpu_100k.err
  run --steps 3 --warmup 3 --no-cpu-baseline --config C3 --batch 12500 > gpurun_out/r02_final_bench_c3_8gpu_100k.json 2> gpurun_out/r02_final_bench_c3_8gpu_100k.err
fi
