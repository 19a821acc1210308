#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/stress_parity.py 2500 55 > gpurun_out/stress_api.txt 2>&1; echo "rc=$?"; tail -6 gpurun_out/stress_api.txt | cut -c1-500
