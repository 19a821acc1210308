#!/bin/bash
# 1 GPU: parity of the refit / LO paths after the eigen-solver change, config times, then the five-point solver at several
# register budgets (resident warps per SM) and solve-ahead depths.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_host_layer.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/gputest_r2g.txt; tail -3 gpurun_out/gputest_r2g.txt
CONFIG_TIMES_NO_CPU=1 CONFIG_TIMES_REPS=5 python tools/config_times.py 2>/dev/null | tee gpurun_out/config_times_r2g.txt
export CONFIG_TIMES_NO_CPU=1 CONFIG_TIMES_REPS=5 CONFIG_TIMES_ONLY="C4 essential N=20000 20% uniform+SPRT (no LO)"
for v in libusac_gpu.so v_e5_3.so v_e5_4.so v_e5_6.so; do
  for g in 2 4 8; do
    echo "== $v solve-ahead $g: $(USAC_GPU_LIB=ransac_b200/$v USAC_GPU_SOLVE_AHEAD=$g USAC_GPU_TRACE=2 python tools/config_times.py 2>gpurun_out/e5_trace.txt | tail -1) $(grep 'kernels (us' gpurun_out/e5_trace.txt | tail -1 | grep -o 'solve=[0-9]*')"
  done
done | tee gpurun_out/e5_ab.txt
