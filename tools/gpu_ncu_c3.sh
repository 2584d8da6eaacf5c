#!/bin/bash
# ncu --set full capture of the C3 lane-per-problem kernel (100k problems, one launch)
tag=${1:-r02_lane_c3}
mkdir -p gpurun_out
python tools/run_case.py C3 --batch 100000 --path fused > gpurun_out/${tag}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_fused_lane -c 1 -o gpurun_out/prof_${tag} -f python tools/run_case.py C3 --batch 100000 --path fused > gpurun_out/${tag}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_plain.log
