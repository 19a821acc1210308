#!/bin/bash
# sweep of the first-round size: RS="128 256 384 512" bash tools/gpu_rs.sh
mkdir -p gpurun_out
for v in $RS; do
  python bench.py --no-cpu --steps 8 --round-size $v 2>&1 | tail -1 > gpurun_out/rs_$v.json
  python - <<PY
import json
d=json.load(open("gpurun_out/rs_$v.json"))
print("round_size=$v", "value %.1f G/s pipelined %.1f e2e %.1f G/s frac %.3f launch_ms %.3f share %.2f useful %.3f launches/step %.0f" % (d["value"]/1e9, d["config"]["value_pipelined"]/1e9, d["e2e"]["value"]/1e9, d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["roofline"]["score_share_of_step"], d["config"]["useful_fraction"], d["gpu_launches"]/d["steps"]))
PY
done
