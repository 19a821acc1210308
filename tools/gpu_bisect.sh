#!/bin/bash
mkdir -p gpurun_out
STRESS_DEBUG=1 STRESS_ONLY_CASES=96 timeout 600 python tools/stress_parity.py 97 21 2>&1 | grep -v "^MISMATCH ('batch" | tail -8 | cut -c1-600
STRESS_DEBUG=1 STRESS_ONLY_CASES=300 timeout 600 python tools/stress_parity.py 301 21 2>&1 | grep "DEBUG" | cut -c1-600
