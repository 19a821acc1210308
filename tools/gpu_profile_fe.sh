#!/bin/bash
# ncu evidence for the epipolar scoring kernels (and a DRAM-streamed first-round launch of the homography kernel):
# plain run first (numbers), then one --set full capture each. Usage: tools/gpu_profile_fe.sh <tag>
TAG=${1:-r2}
mkdir -p gpurun_out
for k in fundamental essential homography; do
  python tools/score_bench.py 1184 $k > gpurun_out/${TAG}_score_bench_$k.txt 2>&1
  tail -3 gpurun_out/${TAG}_score_bench_$k.txt
done
for k in fundamental essential; do
  ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 7 -c 1 -f -o gpurun_out/${TAG}_prof_$k python tools/score_bench.py 1184 $k > gpurun_out/${TAG}_ncu_$k.log 2>&1
  tail -1 gpurun_out/${TAG}_ncu_$k.log
done
python bench.py --no-cpu --no-c5 --steps 1 --warmup 1 > gpurun_out/${TAG}_plain_h.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 5 -c 1 -f -o gpurun_out/${TAG}_prof_h_round1 python bench.py --no-cpu --no-c5 --steps 1 --warmup 1 > gpurun_out/${TAG}_ncu_h.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_h.log
ls -la gpurun_out/${TAG}_prof_*
