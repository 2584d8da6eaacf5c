#!/bin/bash
mkdir -p gpurun_out
python tools/bench_steps.py C3 --only 0,1,2,3,4,5 > gpurun_out/steps_c3.txt 2>&1
python tools/bench_steps.py C2 --only 0,1,2,3,4,5 > gpurun_out/steps_c2.txt 2>&1
python tools/bench_steps.py C2 --batch 100000 --only 0,1,2,3,4,5 > gpurun_out/steps_c2_100k.txt 2>&1
python tools/bench_steps.py C4 > gpurun_out/steps_c4.txt 2>&1
python tools/bench_steps.py C5 --reps 3 > gpurun_out/steps_c5.txt 2>&1
python tools/run_case.py C4 --batch 1000 --reps 2 > gpurun_out/rc_c4.log 2>&1
python tools/run_case.py C5 --reps 2 > gpurun_out/rc_c5.log 2>&1
python tools/run_case.py C5 > gpurun_out/plain_c5.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5.csv python tools/run_case.py C5 > gpurun_out/ncu_c5.log 2>&1
python tools/run_case.py C4 --batch 200 > gpurun_out/plain_c4.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c4.csv python tools/run_case.py C4 --batch 200 > gpurun_out/ncu_c4.log 2>&1
