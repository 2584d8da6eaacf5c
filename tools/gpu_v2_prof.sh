#!/bin/bash
mkdir -p gpurun_out
python tools/run_case.py C3 --batch 3552 --path fused > gpurun_out/plain_c3_v2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_fused2 -c 1 -o gpurun_out/prof_fused2_c3_f -f python tools/run_case.py C3 --batch 3552 --path fused > gpurun_out/ncu_f2_c3.log 2>&1
python tools/run_case.py C2 --batch 592 --path fused > gpurun_out/plain_c2_v2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_fused2 -c 1 -o gpurun_out/prof_fused2_c2_f -f python tools/run_case.py C2 --batch 592 --path fused > gpurun_out/ncu_f2_c2.log 2>&1
