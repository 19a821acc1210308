#!/bin/bash
# final launch lists of round 2 (after the same commands exited 0 without ncu): the default bench step, the C5 fit, one rank of eight
mkdir -p gpurun_out
CMD="python bench.py --no-cpu --no-c5 --no-epipolar --steps 2 --warmup 1"
$CMD > gpurun_out/final_plain.json 2> gpurun_out/final_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_bench_steps2.csv $CMD > /dev/null 2>&1
python bench.py --workload c5 --steps 2 --warmup 1 --no-cpu > gpurun_out/final_plain_c5.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c5.csv python bench.py --workload c5 --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1
python tools/c5_rank_profile.py 8 4 5000 > gpurun_out/c5_rank_plain.txt 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_c5_8gpu.csv python tools/c5_rank_profile.py 8 2 5000 > /dev/null 2>&1
export CONFIG_TIMES_NO_CPU=1 CONFIG_TIMES_REPS=2
CONFIG_TIMES_ONLY="C2 homography N=4000 30% uniform+LO" python tools/config_times.py > /dev/null 2>&1 && \
CONFIG_TIMES_ONLY="C2 homography N=4000 30% uniform+LO" ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c2_lo.csv python tools/config_times.py > /dev/null 2>&1
CONFIG_TIMES_ONLY="C3 fundamental" python tools/config_times.py > /dev/null 2>&1 && \
CONFIG_TIMES_ONLY="C3 fundamental" ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c3.csv python tools/config_times.py > /dev/null 2>&1
ls -la gpurun_out/r2_launches_*.csv; cat gpurun_out/c5_rank_plain.txt | tail -1
