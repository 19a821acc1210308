import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O
from ransac_b200 import GpuContext, capi, generator as gen
lo = int(sys.argv[1]) if len(sys.argv) > 1 else 0
pts = gen.make(4)[0]
ctx = GpuContext(0); ctx.set_points(capi.EST_ESSENTIAL, pts); ctx.set_sprt_pool(0, O.sprt_pool(1, len(pts)))
for _ in range(2):
    r = ctx.fit(2.5e-3, 0.95, 10000, seed=1, round_size=512, sprt=True, lo=lo)[0]
print(r["inliers"], r["iterations"], ctx.last_timing())
