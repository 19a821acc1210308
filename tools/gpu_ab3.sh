#!/bin/bash
# A/B of scoring-kernel build variants on ONE box: bench line (C2) + realistic-data Sampson / essential launches
for lib in "$@"; do
  USAC_GPU_LIB=$PWD/ransac_b200/$lib python bench.py --no-cpu --no-c5 --no-epipolar --steps 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib value %.4e e2e %.4e frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['frac']))"
  SCORE_BENCH_INLIERS=0.25 SCORE_BENCH_K=512,1024 USAC_GPU_LIB=$PWD/ransac_b200/$lib python tools/score_bench.py 1184 fundamental 2>&1 | tail -2 | sed "s/^/$lib /"
  SCORE_BENCH_INLIERS=0.2 SCORE_BENCH_K=512,1024 USAC_GPU_LIB=$PWD/ransac_b200/$lib python tools/score_bench.py 1184 essential 2>&1 | tail -2 | sed "s/^/$lib /"
  SCORE_BENCH_K=512 USAC_GPU_LIB=$PWD/ransac_b200/$lib python tools/score_bench.py 1184 homography 2>&1 | tail -1 | sed "s/^/$lib /"
done
