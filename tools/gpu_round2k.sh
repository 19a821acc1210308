#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_host_layer.py -x -q -m gpu -k "essential or solver or config4 or e5 or five" 2>&1 | tail -8 > gpurun_out/gputest_r2k.txt; tail -3 gpurun_out/gputest_r2k.txt
USAC_GPU_TRACE=2 CONFIG_TIMES_NO_CPU=1 CONFIG_TIMES_REPS=5 CONFIG_TIMES_ONLY="C4" python tools/config_times.py 2>gpurun_out/trace_r2k.txt | tee gpurun_out/config_times_r2k.txt
grep "kernels (us" gpurun_out/trace_r2k.txt | tail -2
