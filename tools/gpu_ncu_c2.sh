#!/bin/bash
# ncu --set full capture of the C2 whole-solve kernel (10k problems, one launch) + the launch list of the bench command
tag=${1:-r02_fused3_c2}
mkdir -p gpurun_out
python tools/run_case.py C2 --batch 10000 --path fused > gpurun_out/${tag}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_fused3 -c 1 -o gpurun_out/prof_${tag} -f python tools/run_case.py C2 --batch 10000 --path fused > gpurun_out/${tag}_ncu.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_plain.json 2> gpurun_out/${tag}_bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
