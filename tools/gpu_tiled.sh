#!/bin/bash
mkdir -p gpurun_out
python tools/run_case.py C4 --batch 1000 --reps 2 > gpurun_out/rc_c4.log 2>&1
python tools/run_case.py C5 --reps 2 > gpurun_out/rc_c5.log 2>&1
python tools/run_case.py C4 --batch 200 > gpurun_out/plain_c4.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c4.csv python tools/run_case.py C4 --batch 200 > gpurun_out/ncu_c4.log 2>&1
python tools/run_case.py C5 > gpurun_out/plain_c5.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5.csv python tools/run_case.py C5 > gpurun_out/ncu_c5.log 2>&1
