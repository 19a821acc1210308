#!/bin/bash
# quick GPU check: parity tests + a short bench (+ optional ncu capture of the scoring kernel when NCU=1)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
python bench.py --no-cpu --steps 5 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; cat gpurun_out/bench_quick.json; tail -5 gpurun_out/bench_quick.err
for rs in $RS; do python bench.py --no-cpu --steps 5 --round-size $rs 2>&1 | tail -1 > gpurun_out/bench_rs$rs.json; done
if [ "$NCU" = "1" ]; then
NCU_CMD="python bench.py --no-cpu --steps 2 --warmup 1 --problems 1184"
$NCU_CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 3 -c 1 -f -o gpurun_out/prof_score $NCU_CMD > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
fi
