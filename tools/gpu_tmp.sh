#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log
python tools/bench_steps.py C3 --only 0,1,2,3,4,5 > gpurun_out/steps_c3.txt 2>&1
python tools/bench_steps.py C2 --batch 100000 --only 0,1,2,3,4,5 > gpurun_out/steps_c2_100k.txt 2>&1
python tools/bench_steps.py C4 --only 0,1,6 > gpurun_out/steps_c4b.txt 2>&1
