import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ransac_b200 import GpuContext, capi, generator as gen
pts = gen.make(5, n=1000000)[0]
host = torch.from_numpy(pts).pin_memory()
ctx = GpuContext(0)
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.set_points(capi.EST_HOMOGRAPHY, host)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    ctx.set_neighbors_grid(0, 50)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"set_points {1e3*(t1-t0):.2f} ms  set_neighbors_grid {1e3*(t2-t1):.2f} ms")
