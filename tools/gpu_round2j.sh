#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_round2e.sh 2
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_host_layer.py tests/test_gpu_multirank.py -x -q -m gpu -k "lo or LO or local or refit or plugin or nccl or peer" 2>&1 | tail -8 > gpurun_out/gputest_r2j.txt; tail -3 gpurun_out/gputest_r2j.txt
CONFIG_TIMES_NO_CPU=1 CONFIG_TIMES_REPS=5 CONFIG_TIMES_ONLY="LO" python tools/config_times.py 2>/dev/null | tee gpurun_out/config_times_r2j.txt
