#!/bin/bash
# One GPU session: parity tests, phase timing, bench, ncu launch list + full capture of the fused kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log
python tools/phase_timing.py C2 --batch 2960 > gpurun_out/phase_c2.txt 2>&1
python tools/phase_timing.py C3 --batch 17760 > gpurun_out/phase_c3.txt 2>&1
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --config C3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
ncu --set full --clock-control none --import-source on -k regex:k_fused -c 1 -o gpurun_out/prof_fused_c2_r1b -f python tools/run_case.py C2 --batch 592 --path fused > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fused -c 1 -o gpurun_out/prof_fused_c3_r1b -f python tools/run_case.py C3 --batch 3552 --path fused > gpurun_out/ncu_full_c3.log 2>&1
