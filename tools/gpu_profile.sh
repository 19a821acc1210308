#!/bin/bash
# ncu evidence for profiles/: launch list of a short bench run + one full capture of the scoring kernel
mkdir -p gpurun_out
NCU_CMD="python bench.py --no-cpu --steps 2 --warmup 1"
$NCU_CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $NCU_CMD > gpurun_out/ncu1.log 2>&1
$NCU_CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 4 -c 1 -f -o gpurun_out/prof_score $NCU_CMD > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log; cat gpurun_out/plain.log | tail -1 | cut -c1-400
