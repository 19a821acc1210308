#!/bin/bash
# 1 GPU: the GPU test suite, ms per fit of the BASELINE configurations against the oracle, the default bench line.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/gputest_check.txt; tail -3 gpurun_out/gputest_check.txt
python tools/config_times.py 2>/dev/null | tee gpurun_out/config_times_check.txt
python bench.py > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_check.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'f',d['roofline_f']['frac'],'e',d['roofline_e']['frac'],'c5',d['config']['c5']['ms_per_fit'])
PY
