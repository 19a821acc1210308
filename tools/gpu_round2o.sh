#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/gputest_r2o.txt; tail -3 gpurun_out/gputest_r2o.txt
python tools/refit_time.py 2>&1 | tail -1 | tee gpurun_out/refit_time.txt
bash tools/gpu_pipe_sweep.sh
