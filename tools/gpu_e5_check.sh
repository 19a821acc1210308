#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_host_layer.py tests/test_gpu_stress.py -x -q -m gpu -k "essential or solver or config4 or stress or randomised or estimate" 2>&1 | tail -6 > gpurun_out/gputest_e5.txt; tail -3 gpurun_out/gputest_e5.txt
USAC_GPU_TRACE=2 CONFIG_TIMES_REPS=5 CONFIG_TIMES_ONLY="C4" python tools/config_times.py 2>gpurun_out/trace_e5.txt | tee gpurun_out/config_times_e5.txt
grep "kernels (us" gpurun_out/trace_e5.txt | tail -1
