#!/bin/bash
# phase-cycle tables of the C2 whole-solve kernel with the -DSOCP_PHASE_TIMING build (tools/phase_timing.py)
tag=${1:-ph}
for b in 148 592 5920; do
  SOCP_B200_LIB=socp.jl_b200/lib/libsocp_b200_prof.so python tools/phase_timing.py C2 --batch $b > gpurun_out/r2_${tag}_phase_c2_$b.txt 2>&1
done
