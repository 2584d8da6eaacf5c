#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log
python bench.py --no-cpu-baseline > gpurun_out/bench_try.json 2> gpurun_out/bench_try.err
python bench.py --config C3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
python tools/phase_timing.py C2 --batch 148 > gpurun_out/phase_c2_fine_148.txt 2>&1
