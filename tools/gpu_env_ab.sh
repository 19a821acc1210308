#!/bin/bash
# A/B on one box over an environment knob: ENVVAR=name VALUES="a b c" bash tools/gpu_env_ab.sh
mkdir -p gpurun_out
for v in $VALUES; do
  env $ENVVAR=$v python bench.py --no-cpu --steps 8 2>&1 | tail -1 > gpurun_out/env_${ENVVAR}_$v.json
  python - <<PY
import json
d=json.load(open("gpurun_out/env_${ENVVAR}_$v.json"))
print("$ENVVAR=$v", "value %.1f G/s pipelined %.1f e2e %.1f G/s frac %.3f launch_ms %.3f share %.2f" % (d["value"]/1e9, d["config"]["value_pipelined"]/1e9, d["e2e"]["value"]/1e9, d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["roofline"]["score_share_of_step"]))
PY
done
