#!/bin/bash
# parity suite + scoring micro-benchmarks + the bench line (no CPU leg). Usage: tools/gpu_quick2.sh <tag>
TAG=${1:-q}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/${TAG}_gputest.txt; cat gpurun_out/${TAG}_gputest.txt
for k in fundamental essential homography; do
  python tools/score_bench.py 1184 $k ${2:-} > gpurun_out/${TAG}_score_bench_$k.txt 2>&1; tail -4 gpurun_out/${TAG}_score_bench_$k.txt
done
python bench.py --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 300 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "c5 ms", d["config"]["c5"]["ms_per_fit"], d["config"]["c5"]["parity_vs_n1"])
PY
