"""Where does the e2e arm lose time? 3 worker threads, each: [upload] -> fit, 10 steps; variants with/without the per-step upload."""
import os, sys, time, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ransac_b200 import GpuContext, capi, generator as gen
B, N, NP = 2368, 4000, 3
pts = np.concatenate([gen.homography(n=N, seed=1000 + i)[0] for i in range(B)])
host = torch.from_numpy(pts).pin_memory()
parts = [(i * B // NP, (i + 1) * B // NP) for i in range(NP)]
ctxs = []
for i in range(NP):
    c = GpuContext(0); s = torch.cuda.Stream(); c.set_stream(s.cuda_stream); c._s = s
    lo, hi = parts[i]; c.set_points(capi.EST_HOMOGRAPHY, host[lo * N:hi * N], [N] * (hi - lo)); ctxs.append(c)
kw = dict(threshold=2.0, confidence=0.95, max_iterations=10000, seed=1, round_size=128)
def run(upload, steps=10):
    tu = [0.0] * NP; tf = [0.0] * NP
    def worker(i):
        lo, hi = parts[i]
        for k in range(steps):
            t0 = time.perf_counter()
            if upload: ctxs[i].set_points(capi.EST_HOMOGRAPHY, host[lo * N:hi * N], [N] * (hi - lo))
            t1 = time.perf_counter()
            ctxs[i].fit_records(**kw)
            t2 = time.perf_counter(); tu[i] += t1 - t0; tf[i] += t2 - t1
    torch.cuda.synchronize(); t0 = time.perf_counter()
    th = [threading.Thread(target=worker, args=(i,)) for i in range(NP)]
    [t.start() for t in th]; [t.join() for t in th]
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"upload={upload}: {dt / steps * 1e3:.2f} ms/step; per thread upload {np.mean(tu) / steps * 1e3:.2f} ms fit {np.mean(tf) / steps * 1e3:.2f} ms")
run(False, 3); run(False); run(True); run(True)
