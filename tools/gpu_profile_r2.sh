#!/bin/bash
# ncu evidence for profiles/ (round 2): launch list of a short bench run, then one --set full capture of the scoring kernel for
# the homography bench (first, DRAM-streamed round of a step), the Sampson and the essential kernel (bench's roofline_f/e launches).
# Every capture follows a plain run of the same command that exited 0.
mkdir -p gpurun_out
CMD="python bench.py --no-cpu --no-c5 --no-epipolar --steps 2 --warmup 1"
$CMD > gpurun_out/r2_plain.json 2> gpurun_out/r2_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench_steps2.csv $CMD > /dev/null 2>&1
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:score_sq_kernel -s 5 -c 1 -f -o gpurun_out/r2_score_sq_h $CMD > gpurun_out/r2_ncu_h.log 2>&1
export SCORE_BENCH_K=1024
SCORE_BENCH_INLIERS=0.25 python tools/score_bench.py 1184 fundamental > gpurun_out/r2_plain_f.txt 2>&1 && \
SCORE_BENCH_INLIERS=0.25 ncu --set full --clock-control none --import-source on -k regex:score_sq_kernel -s 3 -c 1 -f -o gpurun_out/r2_score_sq_f python tools/score_bench.py 1184 fundamental > gpurun_out/r2_ncu_f.log 2>&1
SCORE_BENCH_INLIERS=0.2 python tools/score_bench.py 1184 essential > gpurun_out/r2_plain_e.txt 2>&1 && \
SCORE_BENCH_INLIERS=0.2 ncu --set full --clock-control none --import-source on -k regex:score_sq_kernel -s 3 -c 1 -f -o gpurun_out/r2_score_sq_e python tools/score_bench.py 1184 essential > gpurun_out/r2_ncu_e.log 2>&1
python bench.py --workload c5 --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_plain_c5.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c5.csv python bench.py --workload c5 --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1
tail -1 gpurun_out/r2_plain_f.txt gpurun_out/r2_plain_e.txt; ls -la gpurun_out/r2_score_sq_*.ncu-rep gpurun_out/r2_launches_*.csv
