#!/bin/bash
# 1 GPU: the GPU test suite, the plug-in layer sweep (rounds of 1 and of 4 against the matching oracle forms), smoke, the default bench line.
mkdir -p gpurun_out
make -s -C ransac_b200/usac
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/gputest_final.txt; tail -3 gpurun_out/gputest_final.txt
timeout 300 python tools/stress_harness.py 40 1 1 > gpurun_out/stress_harness_r1.txt 2>&1; echo "rc=$?"; tail -4 gpurun_out/stress_harness_r1.txt | cut -c1-400
timeout 300 python tools/stress_harness.py 30 2 4 > gpurun_out/stress_harness_r4.txt 2>&1; echo "rc=$?"; tail -4 gpurun_out/stress_harness_r4.txt | cut -c1-400
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final2.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'f',d['roofline_f']['frac'],'e',d['roofline_e']['frac'],'c5',d['config']['c5']['ms_per_fit'])
PY
