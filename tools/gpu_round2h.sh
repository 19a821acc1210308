#!/bin/bash
# 1 GPU: LO parity tests, config times, source-level capture of lo_kernel on the 20000-point essential problem (C4 + LO).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_host_layer.py -x -q -m gpu -k "lo or LO or local or refit or plugin" 2>&1 | tail -8 > gpurun_out/gputest_r2h.txt; tail -3 gpurun_out/gputest_r2h.txt
CONFIG_TIMES_NO_CPU=1 CONFIG_TIMES_REPS=5 python tools/config_times.py 2>/dev/null | tee gpurun_out/config_times_r2h.txt
export CONFIG_TIMES_NO_CPU=1 CONFIG_TIMES_REPS=1
CONFIG_TIMES_ONLY="C4 essential N=20000 20% uniform+SPRT+LO" ncu --set full --clock-control none --import-source on -k regex:lo_kernel -c 1 -f -o gpurun_out/r2_lo_kernel_c4 python tools/config_times.py > gpurun_out/r2_ncu_lo_c4.log 2>&1
tail -2 gpurun_out/r2_ncu_lo_c4.log
