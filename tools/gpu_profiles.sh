#!/bin/bash
# Round-end evidence: bench lines, ncu launch list of the bench command, ncu --set full of the dominant kernels.
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python tools/run_case.py C2 --batch 10000 --path fused > gpurun_out/plain_c2_10k.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_fused2 -c 1 -o gpurun_out/prof_r01_fused_c2_10k -f python tools/run_case.py C2 --batch 10000 --path fused > gpurun_out/ncu_c2_10k.log 2>&1
python tools/run_case.py C3 --batch 100000 --path fused > gpurun_out/plain_c3_100k.log 2>&1 && \
ncu --set full --clock-control none -k regex:k_fused2 -c 1 -o gpurun_out/prof_r01_fused_c3_100k -f python tools/run_case.py C3 --batch 100000 --path fused > gpurun_out/ncu_c3_100k.log 2>&1
python tools/bench_steps.py C5 --reps 2 --only 7,8 > gpurun_out/plain_steps_c5.log 2>&1 && \
ncu --set full --clock-control none -k regex:k_syrk -c 2 -o gpurun_out/prof_r01_syrk_c5 -f python tools/bench_steps.py C5 --reps 1 --only 7 > gpurun_out/ncu_syrk_c5.log 2>&1
python tools/bench_steps.py C3 --reps 2 --only 0,1,3 > gpurun_out/plain_steps_c3.log 2>&1 && \
ncu --set full --clock-control none -k regex:bk_ -c 6 -o gpurun_out/prof_r01_cone_c3 -f python tools/bench_steps.py C3 --reps 1 --only 0,1,3 > gpurun_out/ncu_cone_c3.log 2>&1
