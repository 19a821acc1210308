#!/bin/bash
# kernel-tuning sweep: the same short bench against every variant library listed in $VARIANTS
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
for v in $VARIANTS; do
  USAC_GPU_LIB=$PWD/ransac_b200/$v python bench.py --no-cpu --steps 5 2>&1 | tail -1 > gpurun_out/bench_$v.json
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$v.json"))
print("$v", "value %.1f G/s e2e %.1f G/s frac %.3f launch_ms %.3f share %.2f" % (d["value"]/1e9, d["e2e"]["value"]/1e9, d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["roofline"]["score_share_of_step"]))
PY
done
