#!/bin/bash
mkdir -p gpurun_out
make -s -C ransac_b200/usac
timeout 900 python tools/stress_harness.py 60 1 > gpurun_out/stress_harness.txt 2>&1; echo "rc=$?"; tail -8 gpurun_out/stress_harness.txt | cut -c1-400
