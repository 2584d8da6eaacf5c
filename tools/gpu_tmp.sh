#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --config C3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python tools/run_case.py C4 --batch 200 > gpurun_out/plain_c4.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c4.csv python tools/run_case.py C4 --batch 200 > gpurun_out/ncu_c4.log 2>&1
