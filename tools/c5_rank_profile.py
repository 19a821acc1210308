#!/usr/bin/env python
"""What ONE rank of an R-rank hypothesis-sharded C5 fit executes, on one GPU: rank 0 of R with the exchange replaced by a device-side
copy of this rank's own part into every slot (so select_kernel sees R copies of rank 0's scores - timing only, the result is not
the real fit's). ncu cannot profile a multi-rank command; this gives the per-rank launch list (profiles/r2_launches_c5_8gpu.csv).
usage: c5_rank_profile.py [ranks=8] [fits=5] [K=5000]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ransac_b200 import GpuContext, capi  # noqa: E402
from ransac_b200 import generator as gen  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 8
FITS = int(sys.argv[2]) if len(sys.argv) > 2 else 5
K = int(sys.argv[3]) if len(sys.argv) > 3 else 5000

rt = None
for name in ("libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so.12", "libcudart.so"):
    try:
        rt = C.CDLL(name)
        break
    except OSError:
        continue
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]


def hook(user, d_send, d_recv, nbytes, stream):
    for r in range(R):
        if rt.cudaMemcpyAsync(d_recv + r * nbytes, d_send, nbytes, 3, stream) != 0:      # device -> device, stream ordered
            return 1
    return 0


fn = capi.ALLGATHER_FN(hook)
pts = gen.make(5)[0]
ctx = GpuContext(0)
ctx.set_points(capi.EST_HOMOGRAPHY, pts)
ctx.set_neighbors_grid(0, 50)
if R > 1:
    ctx._check(ctx.L.usac_gpu_set_allgather(ctx.h, fn, None), "set_allgather")
kw = dict(threshold=2.0, confidence=0.95, max_iterations=10000, seed=1, round_size=K, sampler=capi.SAMPLER_NAPSAC, neighbors=capi.NEIGH_GRID,
          rank=0, nranks=R)
times = []
for i in range(FITS + 2):
    t0 = time.perf_counter()
    r = ctx.fit_records(**kw)[0]
    times.append((time.perf_counter() - t0) * 1e3)
    t = ctx.last_timing()
print(f"rank 0 of {R}: K={K} wall ms/fit median {np.median(times[2:]):.3f} min {min(times[2:]):.3f}; GPU events total {t['total_ms']:.3f} ms, scoring {t['score_ms']:.3f} ms, "
      f"launches {t['launches']}, iterations {int(r['iterations'])} rounds {int(r['rounds'])}")
ctx.close()
