#!/bin/bash
# 1 GPU, end of round: the GPU test suite, the parity sweep that found the out-of-reach bug (seed 78), ms per fit of the configurations, smoke, the default bench line
mkdir -p gpurun_out
make -s -C ransac_b200/usac
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/gputest_final.txt; tail -3 gpurun_out/gputest_final.txt
timeout 300 python tools/stress_parity.py 800 78 > gpurun_out/stress78_fixed.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/stress78_fixed.txt | cut -c1-500
python tools/config_times.py 2>/dev/null | tee gpurun_out/config_times_final.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_final3.json 2> gpurun_out/bench_final3.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final3.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'f',d['roofline_f']['frac'],'e',d['roofline_e']['frac'],'c5',d['config']['c5']['ms_per_fit'])
PY
