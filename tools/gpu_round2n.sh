#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sprt" 2>&1 | tail -8 > gpurun_out/gputest_r2n.txt; tail -3 gpurun_out/gputest_r2n.txt
for c in 3 2; do for b in 1 0; do USAC_GPU_SPRT_BATCH=$b python tools/sprt_batch_time.py 256 $c 4000 2>&1 | tail -1; done; done | tee gpurun_out/sprt_batch.txt
