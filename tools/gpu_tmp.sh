#!/bin/bash
bash tools/gpu_round.sh
python tools/phase_timing.py C2 --batch 148 > gpurun_out/phase_c2_fine_148.txt 2>&1
python tools/phase_timing.py C2 --batch 2960 > gpurun_out/phase_c2_fine_2960.txt 2>&1
python tools/phase_timing.py C3 --batch 17760 > gpurun_out/phase_c3_fine.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
