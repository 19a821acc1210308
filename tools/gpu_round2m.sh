#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/gputest_r2m.txt; tail -3 gpurun_out/gputest_r2m.txt
python tools/config_times.py 2>/dev/null | tee gpurun_out/config_times_r2m.txt
python bench.py > gpurun_out/bench_r2m.json 2> gpurun_out/bench_r2m.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2m.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'f',d['roofline_f']['frac'],'e',d['roofline_e']['frac'],'c5',d['config']['c5']['ms_per_fit'])
PY
