#!/bin/bash
# the two mismatches of sweep seed 78: the exact-chain build (round-2 code before the log-domain tail) and the current build, same sweep
mkdir -p gpurun_out
STRESS_DEBUG=1 USAC_GPU_LIB=$PWD/ransac_b200/v_exactwalk.so timeout 300 python tools/stress_parity.py 800 78 > gpurun_out/stress78_exactwalk.txt 2>&1; echo "rc=$?"; grep -v "^  " gpurun_out/stress78_exactwalk.txt | tail -12 | cut -c1-700
STRESS_DEBUG=1 timeout 300 python tools/stress_parity.py 800 78 > gpurun_out/stress78_current.txt 2>&1; echo "rc=$?"; tail -12 gpurun_out/stress78_current.txt | cut -c1-700
