#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_host_layer.py tests/test_gpu_stress.py -x -q -m gpu -k "sprt or config3 or config4 or stress or randomised or plugin" 2>&1 | tail -6 > gpurun_out/gputest_sprt.txt; tail -3 gpurun_out/gputest_sprt.txt
CONFIG_TIMES_REPS=7 CONFIG_TIMES_ONLY="SPRT" python tools/config_times.py 2>gpurun_out/trace_sprt.txt | tee gpurun_out/config_times_sprt.txt
USAC_GPU_SOLVE_OVERLAP=0 CONFIG_TIMES_NO_CPU=1 CONFIG_TIMES_REPS=7 CONFIG_TIMES_ONLY="C4" python tools/config_times.py 2>/dev/null | sed "s/^/no overlap: /"
grep "kernels (us" gpurun_out/trace_sprt.txt | head -3 | tail -1
