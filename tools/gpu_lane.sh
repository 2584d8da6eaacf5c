#!/bin/bash
# lane-per-problem kernel (fused_lane.cuh): parity tests, then the C3 bench line per variant (problems per SM, lanes
# per warp, ring depth) and of fused_v2's one-warp teams
tag=${1:-r02_lane}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "lane or c3_full" > gpurun_out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_tests.log
tail -5 gpurun_out/${tag}_tests.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --config C3 --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/${tag}_bench_c3_${name}.json 2> gpurun_out/${tag}_bench_c3_${name}.err
  echo "$name $(cut -c40-160 gpurun_out/${tag}_bench_c3_${name}.json)"
}
run pps96_lpw32 A=1
run pps96_lpw16 SOCP_B200_LANE_LPW=16
run pps64_lpw32 SOCP_B200_LANE_PPS=64
run pps64_lpw16 SOCP_B200_LANE_PPS=64 SOCP_B200_LANE_LPW=16
run pps64_lpw8 SOCP_B200_LANE_PPS=64 SOCP_B200_LANE_LPW=8
run rs2_pps64_lpw32 SOCP_B200_LANE_RS2=1
run v2 SOCP_B200_LANE=0
