#!/bin/bash
# One GPU session of round-end evidence: parity tests, bench lines of every BASELINE config, the reference arm, the
# ncu --set full capture of the dominant kernel and the launch list of the bench command.  Outputs in gpurun_out/.
tag=${1:-r02_final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_tests.log
python bench.py > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_c2_reference.json 2> gpurun_out/${tag}_bench_c2_reference.err
for cfg in C3 C4 C5; do
  python bench.py --config $cfg --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/${tag}_bench_${cfg}.json 2> gpurun_out/${tag}_bench_${cfg}.err
done
bash tools/gpu_ncu_c2.sh ${tag}_fused3_c2
bash tools/gpu_ncu_c3.sh ${tag}_lane_c3
python tools/bench_steps.py C4 --reps 3 > gpurun_out/${tag}_steps_C4.txt 2>&1
python tools/bench_steps.py C5 --reps 3 > gpurun_out/${tag}_steps_C5.txt 2>&1
python tools/bench_steps.py C2 --batch 100000 --reps 3 > gpurun_out/${tag}_steps_C2_100k.txt 2>&1
python tools/bench_steps.py C3 --reps 3 > gpurun_out/${tag}_steps_C3.txt 2>&1
