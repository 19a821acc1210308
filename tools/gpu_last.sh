#!/bin/bash
# 1 GPU: one more parity sweep (new seed) and the C3 launch list after the log-domain SPRT tail
mkdir -p gpurun_out
STRESS_DEBUG=1 timeout 300 python tools/stress_parity.py 1500 79 > gpurun_out/stress79.txt 2>&1; echo "rc=$?"; tail -5 gpurun_out/stress79.txt | cut -c1-600
export CONFIG_TIMES_NO_CPU=1 CONFIG_TIMES_REPS=2
CONFIG_TIMES_ONLY="C3 fundamental" python tools/config_times.py > /dev/null 2>&1 && \
CONFIG_TIMES_ONLY="C3 fundamental" ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c3_logwalk.csv python tools/config_times.py > /dev/null 2>&1
ls -la gpurun_out/r2_launches_c3_logwalk.csv
