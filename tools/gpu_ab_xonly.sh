#!/bin/bash
# A/B: homography rejection test on one coordinate (v_xonly.so) against the two-coordinate test (libusac_gpu.so), same box.
mkdir -p gpurun_out
for v in libusac_gpu.so v_xonly.so; do
  export USAC_GPU_LIB=ransac_b200/$v
  for rep in 1 2; do
    python bench.py --no-cpu --no-c5 --no-epipolar > gpurun_out/ab_x.json 2>/dev/null
    python - $v $rep <<'PY'
import json,sys
d=json.loads(open('gpurun_out/ab_x.json').read().strip().splitlines()[-1])
print(sys.argv[1],'rep',sys.argv[2],'value %.4e'%d['value'],'e2e %.4e'%d['e2e']['value'],'frac %.3f'%d['roofline']['frac'],'useful',round(d['config']['useful_fraction'],4))
PY
  done
  python tools/score_bench.py 1184 homography 2>&1 | tail -1 | sed "s/^/$v /"
  SCORE_BENCH_INLIERS=0.3 python tools/score_bench.py 1184 homography 2>&1 | tail -1 | sed "s/^/$v (30% inliers) /"
  python tools/c5_rank_profile.py 1 4 5000 2>/dev/null | tail -1 | sed "s/^/$v /"
done | tee gpurun_out/ab_xonly.txt
USAC_GPU_LIB=ransac_b200/v_xonly.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py -x -q -m gpu 2>&1 | tail -3
