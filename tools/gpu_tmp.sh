#!/bin/bash
mkdir -p gpurun_out
for v in A B; do
SOCP_B200_LIB=$PWD/socp.jl_b200/lib/libsocp_b200_$v.so python bench.py --no-cpu-baseline > gpurun_out/bench_var$v.json 2> gpurun_out/bench_var$v.err
SOCP_B200_LIB=$PWD/socp.jl_b200/lib/libsocp_b200_$v.so python bench.py --config C3 --no-cpu-baseline > gpurun_out/bench_c3_var$v.json 2> gpurun_out/bench_c3_var$v.err
done
