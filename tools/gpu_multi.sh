#!/bin/bash
# N GPUs (default 2): the torchrun NCCL + peer-window parity check, then the default bench at N.
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29731 tests/multi_gpu_check.py > gpurun_out/multi_gpu_check_n$N.txt 2>&1; echo "check rc=$?"; tail -3 gpurun_out/multi_gpu_check_n$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus $N --no-epipolar > gpurun_out/bench_multi_n$N.json 2> gpurun_out/bench_multi_n$N.err; echo "bench rc=$?"
python - $N <<'PY'
import json,sys
n=sys.argv[1]
d=json.loads(open(f'gpurun_out/bench_multi_n{n}.json').read().strip().splitlines()[-1])
c=d['config']['c5']
print('value',d['value'],'e2e',d['e2e']['value'],'h2d_only_ms',d['e2e'].get('h2d_only_ms_per_step'),'GB/s/gpu',d['e2e'].get('h2d_only_gb_per_s_per_gpu'),'c5',{k:c[k] for k in ('exchange','ms_per_fit_nccl','ms_per_fit','parity_vs_n1','ranks_agree','gpu_launches_per_fit')})
PY
tail -3 gpurun_out/bench_multi_n$N.err
