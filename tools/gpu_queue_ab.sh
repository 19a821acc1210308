#!/bin/bash
# Next-round starting point for the survivor-queue scoring kernel (score_queue.cuh):
#   parity of variant 2, then bench + score_bench for USAC_GPU_SCORE_QUEUE = 0 (default kernel), 1 (inlined drain), 2 (out-of-line
#   drain) on one box, then one ncu --set full capture of the queue kernel (NCU=1).
mkdir -p gpurun_out
USAC_GPU_SCORE_QUEUE=2 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_queue2.log 2>&1; echo "variant 2 pytest rc=$?"; tail -3 gpurun_out/pytest_queue2.log
for v in 0 1 2; do
  USAC_GPU_SCORE_QUEUE=$v python bench.py --no-cpu --steps 8 2>/dev/null | tail -1 > gpurun_out/queue_$v.json
  python - <<PY
import json
d=json.load(open("gpurun_out/queue_$v.json"))
print("queue=$v", "value %.1f G/s e2e %.1f G/s frac %.3f launch_ms %.3f share %.2f" % (d["value"]/1e9, d["e2e"]["value"]/1e9, d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["roofline"]["score_share_of_step"]))
PY
  USAC_GPU_SCORE_QUEUE=$v python tools/score_bench.py 2368 homography 2>&1 | head -2
done
if [ "$NCU" = "1" ]; then
  CMD="python bench.py --no-cpu --steps 2 --warmup 1"
  USAC_GPU_SCORE_QUEUE=${QV:-2} ncu --set full --clock-control none --import-source on -k regex:score_queue_kernel -s 1 -c 1 -f -o gpurun_out/prof_queue $CMD > gpurun_out/ncu_queue.log 2>&1
  tail -2 gpurun_out/ncu_queue.log
fi
