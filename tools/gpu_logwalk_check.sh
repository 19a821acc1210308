#!/bin/bash
# 1 GPU: the log-domain SPRT tail (USAC_SPRT_LOGWALK) - SPRT parity tests, C3 / C4 times against the exact-chain build, a parity sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "sprt or config3 or config4 or local_optimisation or plugin_classes or stress" 2>&1 | tail -4 > gpurun_out/gputest_logwalk.txt; tail -3 gpurun_out/gputest_logwalk.txt
CONFIG_TIMES_ONLY="SPRT" CONFIG_TIMES_NO_CPU=1 python tools/config_times.py 2>/dev/null | tee gpurun_out/config_times_logwalk2.txt
timeout 400 python tools/stress_parity.py 800 78 > gpurun_out/stress_logwalk2.txt 2>&1; echo "rc=$?"; tail -4 gpurun_out/stress_logwalk2.txt | cut -c1-500
