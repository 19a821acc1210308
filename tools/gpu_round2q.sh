#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/stress_parity.py 400 7 > gpurun_out/stress_parity.txt 2>&1; echo "stress rc=$?"; tail -12 gpurun_out/stress_parity.txt
