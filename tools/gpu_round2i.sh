#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_host_layer.py -x -q -m gpu -k "lo or LO or local or refit or plugin" 2>&1 | tail -8 > gpurun_out/gputest_r2i.txt; tail -3 gpurun_out/gputest_r2i.txt
CONFIG_TIMES_NO_CPU=1 CONFIG_TIMES_REPS=5 python tools/config_times.py 2>/dev/null | tee gpurun_out/config_times_r2i.txt
