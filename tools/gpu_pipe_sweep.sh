#!/bin/bash
mkdir -p gpurun_out
for p in 1 2 4; do
  python bench.py --no-cpu --no-c5 --no-epipolar --pipe $p > gpurun_out/pipe_$p.json 2>/dev/null
  python - $p <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/pipe_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('pipe',sys.argv[1],'value %.4e'%d['value'],'e2e %.4e'%d['e2e']['value'],'e2e ms/step %.3f'%d['e2e']['ms_per_step'],'h2d_only %.3f ms'%d['e2e']['h2d_only_ms_per_step'],'%.1f GB/s'%d['e2e']['h2d_only_gb_per_s_per_gpu'])
PY
done | tee gpurun_out/pipe_sweep.txt
