#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/tests_v2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_v2.log
python tools/phase_timing.py C2 --batch 2960 > gpurun_out/phase_c2_v2.txt 2>&1
python tools/phase_timing.py C3 --batch 17760 > gpurun_out/phase_c3_v2.txt 2>&1
python tools/phase_timing.py C2 --batch 148 > gpurun_out/phase_c2_v2_148.txt 2>&1
