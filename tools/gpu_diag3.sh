#!/bin/bash
# round-2 diagnostics: default bench line, host/device time split of the replay-path configurations, the per-rank view of an
# 8-rank C5 fit (emulated on one GPU) with its launch list, and a source-level capture of lo_kernel.
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo "bench rc=$?"
USAC_GPU_TRACE=2 CONFIG_TIMES_NO_CPU=1 CONFIG_TIMES_REPS=4 python tools/config_times.py > gpurun_out/config_times_trace.txt 2> gpurun_out/trace_configs.txt; echo "config rc=$?"
for R in 1 8; do USAC_GPU_TRACE=2 python tools/c5_rank_profile.py $R 4 5000 > gpurun_out/c5_rank_r$R.txt 2> gpurun_out/trace_c5_r$R.txt; done
python tools/c5_rank_profile.py 8 4 5000 > gpurun_out/c5_rank_plain.txt 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_c5_8gpu.csv python tools/c5_rank_profile.py 8 2 5000 > /dev/null 2>&1
CONFIG_TIMES_ONLY="C2 homography N=4000 30% uniform+LO" CONFIG_TIMES_NO_CPU=1 CONFIG_TIMES_REPS=3 python tools/config_times.py > gpurun_out/lo_plain.txt 2>&1 && \
CONFIG_TIMES_ONLY="C2 homography N=4000 30% uniform+LO" CONFIG_TIMES_NO_CPU=1 CONFIG_TIMES_REPS=1 ncu --set full --clock-control none --import-source on -k regex:lo_kernel -s 2 -c 1 -f -o gpurun_out/r2_lo_kernel python tools/config_times.py > gpurun_out/r2_ncu_lo.log 2>&1
cat gpurun_out/bench_r2c.json | head -c 3000; echo; cat gpurun_out/config_times_trace.txt gpurun_out/c5_rank_r1.txt gpurun_out/c5_rank_r8.txt
ls -la gpurun_out/r2_lo_kernel.ncu-rep gpurun_out/r2_launches_c5_8gpu.csv
