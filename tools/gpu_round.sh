#!/bin/bash
# One gpurun call: parity tests, smoke, bench, then the ncu launch list + one full capture of the scoring kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
python bench.py --latency > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
for rs in 64 256 512; do python bench.py --no-cpu --steps 5 --round-size $rs 2>&1 | tail -1 > gpurun_out/bench_rs$rs.json; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
NCU_CMD="python bench.py --no-cpu --steps 2 --warmup 1 --problems 592"
$NCU_CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $NCU_CMD > gpurun_out/ncu1.log 2>&1
$NCU_CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 3 -c 2 -o gpurun_out/prof_score $NCU_CMD > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
