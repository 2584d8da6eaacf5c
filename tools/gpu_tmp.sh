#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log
python tools/bench_steps.py C5 --reps 3 --only 7,8 > gpurun_out/steps_c5b.txt 2>&1
python tools/bench_steps.py C4 --batch 64 --only 7,8 > gpurun_out/steps_c4_64.txt 2>&1
python tools/run_case.py C5 --reps 2 > gpurun_out/rc_c5.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
