#!/bin/bash
# 1 GPU: the GPU test suite, then the host/device split of the replay-path configurations.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/gputest_r2d.txt; tail -6 gpurun_out/gputest_r2d.txt
USAC_GPU_TRACE=2 CONFIG_TIMES_NO_CPU=1 CONFIG_TIMES_REPS=4 python tools/config_times.py > gpurun_out/config_times_r2d.txt 2> gpurun_out/trace_configs_r2d.txt; echo "config rc=$?"
cat gpurun_out/config_times_r2d.txt
