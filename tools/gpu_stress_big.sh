#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "napsac" 2>&1 | tail -4
timeout 1500 python tools/stress_parity.py 1500 21 > gpurun_out/stress_big.txt 2>&1; echo "rc=$?"; tail -4 gpurun_out/stress_big.txt | cut -c1-400
USAC_GPU_LO_SEQ=1 USAC_GPU_SPRT_BATCH=0 USAC_GPU_SOLVE_OVERLAP=0 timeout 900 python tools/stress_parity.py 300 22 > gpurun_out/stress_big_fallback.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/stress_big_fallback.txt | cut -c1-400
