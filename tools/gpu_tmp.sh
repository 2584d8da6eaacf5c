#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log
python tools/bench_steps.py C4 > gpurun_out/steps_c4.txt 2>&1
python tools/bench_steps.py C5 --reps 3 > gpurun_out/steps_c5.txt 2>&1
python tools/run_case.py C5 --reps 2 > gpurun_out/rc_c5.log 2>&1
