#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_host_layer.py -x -q -m gpu -k "refit or nonminimal or lo or LO or local or plugin or fused or harness" 2>&1 | tail -6 > gpurun_out/gputest_r2p.txt; tail -3 gpurun_out/gputest_r2p.txt
python tools/refit_time.py 2>&1 | tail -1 | tee gpurun_out/refit_time.txt
