#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log
python tools/bench_steps.py C4 --only 6 > gpurun_out/steps_c4b.txt 2>&1
python tools/bench_steps.py C5 --reps 3 --only 6 > gpurun_out/steps_c5b.txt 2>&1
python tools/bench_steps.py C2 --batch 100000 --only 6 > gpurun_out/steps_c2b.txt 2>&1
python tools/run_case.py C4 --batch 1000 --reps 2 > gpurun_out/rc_c4.log 2>&1
