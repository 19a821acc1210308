#!/bin/bash
# A/B on one box: the short bench against each library in $VARIANTS, in the order given (repeat names to alternate)
mkdir -p gpurun_out
i=0
for v in $VARIANTS; do
  i=$((i+1))
  USAC_GPU_LIB=$PWD/ransac_b200/$v python bench.py --no-cpu --steps 8 2>&1 | tail -1 > gpurun_out/ab_${i}_$v.json
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_${i}_$v.json"))
print("$v", "value %.1f G/s pipelined %.1f e2e %.1f G/s frac %.3f launch_ms %.3f share %.2f" % (d["value"]/1e9, d["config"]["value_pipelined"]/1e9, d["e2e"]["value"]/1e9, d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["roofline"]["score_share_of_step"]))
PY
done
