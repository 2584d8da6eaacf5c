#!/bin/bash
mkdir -p gpurun_out
python tools/run_case.py C2 --batch 10000 --reps 3 > gpurun_out/nw8_10k.log 2>&1
SOCP_B200_F2_NW4=1 python tools/run_case.py C2 --batch 10000 --reps 3 > gpurun_out/nw4_10k.log 2>&1
python tools/run_case.py C2 --batch 296 --reps 3 > gpurun_out/nw8_296.log 2>&1
SOCP_B200_F2_NW4=1 python tools/run_case.py C2 --batch 296 --reps 3 > gpurun_out/nw4_296.log 2>&1
python tools/run_case.py C2 --batch 148 --reps 3 > gpurun_out/nw8_148.log 2>&1
SOCP_B200_F2_NW4=1 python tools/run_case.py C2 --batch 148 --reps 3 > gpurun_out/nw4_148.log 2>&1
