#!/bin/bash
mkdir -p gpurun_out
for g in 0 2 4 16; do for rs in 128 256; do
  USAC_GPU_ROUND_GROWTH=$g python bench.py --no-cpu --steps 5 --round-size $rs 2>&1 | tail -1 > gpurun_out/bench_g${g}_rs$rs.json
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_g${g}_rs$rs.json"))
print("growth $g rs $rs", "value %.1f G/s e2e %.1f G/s frac %.3f launch_ms %.3f share %.2f step_ms %.2f launches %d useful %.2f" % (d["value"]/1e9, d["e2e"]["value"]/1e9, d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["roofline"]["score_share_of_step"], d["ms_per_step"], d["gpu_launches"], d["config"]["useful_fraction"]))
PY
done; done
