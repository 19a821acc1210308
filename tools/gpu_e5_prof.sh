#!/bin/bash
mkdir -p gpurun_out
export CONFIG_TIMES_NO_CPU=1 CONFIG_TIMES_REPS=1
CONFIG_TIMES_ONLY="C4 essential N=20000 20% uniform+SPRT (no LO)" ncu --set full --clock-control none --import-source on -k regex:e5_warp -c 1 -f -o gpurun_out/r2_e5_solver_c python tools/config_times.py > gpurun_out/r2_ncu_e5c.log 2>&1
ls -la gpurun_out/r2_e5_solver_c.ncu-rep
