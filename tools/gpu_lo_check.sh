#!/bin/bash
# LO-RANSAC: parity tests + randomised sweep, then ms per fit with the speculative waves and with the sequential kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_host_layer.py tests/test_gpu_stress.py -x -q -m gpu -k "lo or LO or local or plugin or stress or randomised" 2>&1 | tail -6 > gpurun_out/gputest_lo.txt; tail -3 gpurun_out/gputest_lo.txt
export CONFIG_TIMES_NO_CPU=1 CONFIG_TIMES_REPS=5 CONFIG_TIMES_ONLY="LO"
USAC_GPU_TRACE=1 python tools/config_times.py 2>gpurun_out/lo_trace.txt | tee gpurun_out/lo_times.txt
grep "replay\]" gpurun_out/lo_trace.txt | tail -2 | cut -c1-260
USAC_GPU_LO_SEQ=1 python tools/config_times.py 2>/dev/null | sed 's/^/sequential kernel: /' | tee -a gpurun_out/lo_times.txt
