#!/bin/bash
# lane-per-problem kernel (fused_lane.cuh): parity tests, then the C3 bench line per lanes-per-warp variant
tag=${1:-r02_lane}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "lane or c3_full" > gpurun_out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_tests.log
tail -5 gpurun_out/${tag}_tests.log
for lpw in 8 16 32 4; do
  SOCP_B200_LANE_LPW=$lpw timeout 300 python bench.py --config C3 --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/${tag}_bench_c3_lpw${lpw}.json 2> gpurun_out/${tag}_bench_c3_lpw${lpw}.err
  echo "lpw=$lpw $(cut -c1-160 gpurun_out/${tag}_bench_c3_lpw${lpw}.json)"
done
SOCP_B200_LANE=0 timeout 300 python bench.py --config C3 --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/${tag}_bench_c3_v2.json 2>/dev/null
echo "v2 $(cut -c1-160 gpurun_out/${tag}_bench_c3_v2.json)"
