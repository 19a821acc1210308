#!/bin/bash
# work-item size of the scoring kernel for one 1M-point fit, as one rank of 1 / 2 / 4 / 8 (emulated on one GPU)
mkdir -p gpurun_out
for R in 1 2 4 8; do for T in 2 4 8; do
  echo "R=$R min_tiles=$T: $(USAC_GPU_MIN_TILES=$T python tools/c5_rank_profile.py $R 6 5000 2>/dev/null | tail -1)"
done; done | tee gpurun_out/tiles_sweep.txt
