#!/bin/bash
# 1 GPU: the GPU test suite; then source-level captures of lo_kernel (C2 + LO) and the five-point solver (C4).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/gputest_r2f.txt; tail -4 gpurun_out/gputest_r2f.txt
export CONFIG_TIMES_NO_CPU=1 CONFIG_TIMES_REPS=1
CONFIG_TIMES_ONLY="C2 homography N=4000 30% uniform+LO" ncu --set full --clock-control none --import-source on -k regex:lo_kernel -c 1 -f -o gpurun_out/r2_lo_kernel python tools/config_times.py > gpurun_out/r2_ncu_lo.log 2>&1
tail -2 gpurun_out/r2_ncu_lo.log
CONFIG_TIMES_ONLY="C4 essential N=20000 20% uniform+SPRT (no LO)" ncu --set full --clock-control none --import-source on -k regex:e5_warp -c 1 -f -o gpurun_out/r2_e5_solver python tools/config_times.py > gpurun_out/r2_ncu_e5.log 2>&1
tail -2 gpurun_out/r2_ncu_e5.log
ls -la gpurun_out/*.ncu-rep
