#!/bin/bash
# 1 GPU: the plug-in layer sweep with NAPSAC cases (rounds of 1 and of 8)
mkdir -p gpurun_out
make -s -C ransac_b200/usac
timeout 400 python tools/stress_harness.py 40 3 1 napsac > gpurun_out/stress_harness_napsac_r1.txt 2>&1; echo "rc=$?"; tail -6 gpurun_out/stress_harness_napsac_r1.txt | cut -c1-600
timeout 400 python tools/stress_harness.py 30 4 8 napsac > gpurun_out/stress_harness_napsac_r8.txt 2>&1; echo "rc=$?"; tail -6 gpurun_out/stress_harness_napsac_r8.txt | cut -c1-600
