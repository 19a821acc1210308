#!/bin/bash
# sensitivity of the small-launch scoring benchmarks to the number of work items per warp (point-axis chunking)
for ipw in 16 64 256 1024; do
  for k in fundamental essential; do
    echo "== items/warp $ipw $k"; USAC_GPU_ITEMS_PER_WARP=$ipw python tools/score_bench.py 1184 $k 2>&1 | tail -3
  done
done
