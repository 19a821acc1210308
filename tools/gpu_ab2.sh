#!/bin/bash
# A/B of scoring-kernel variants on ONE box: for every library, the bench line (no CPU / c5 legs) and the small-launch benches
mkdir -p gpurun_out
for lib in "$@"; do
  for rep in 1 2; do
    USAC_GPU_LIB=$PWD/ransac_b200/$lib python bench.py --no-cpu --no-c5 --steps 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib rep$rep value %.4e e2e %.4e frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['frac']))"
  done
  for k in fundamental essential homography; do USAC_GPU_LIB=$PWD/ransac_b200/$lib python tools/score_bench.py 1184 $k 2>&1 | tail -1 | sed "s/^/$lib /"; done
done
echo "== legacy kernel (USAC_GPU_SCORE_LEGACY=1)"
USAC_GPU_SCORE_LEGACY=1 python bench.py --no-cpu --no-c5 --steps 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('legacy value %.4e e2e %.4e frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['frac']))"
for k in fundamental essential homography; do USAC_GPU_SCORE_LEGACY=1 python tools/score_bench.py 1184 $k 2>&1 | tail -1 | sed "s/^/legacy /"; done
