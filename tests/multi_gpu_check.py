"""torchrun entry (one process per GPU): hypothesis-sharded fit of one large homography problem over NCCL; every rank must
produce the oracle's result. Launched by tests/test_gpu_multirank.py::test_nccl_two_gpus and usable by hand:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from oracle import oracle as O
    from ransac_b200 import GpuContext, nccl_unique_id
    from ransac_b200 import dist as D
    from ransac_b200 import generator as gen
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    rank, world = D.init("nccl")
    uid = D.broadcast_bytes(nccl_unique_id() if rank == 0 else None)
    pts = gen.homography(n=200000, inlier_ratio=0.2, seed=77)[0]
    ctx = GpuContext(local)
    ctx.set_points(O.EST_HOMOGRAPHY, pts)
    ctx.nccl_init(uid, rank, world)
    r = ctx.fit(2.0, 0.95, 4000, seed=3, round_size=512, rank=rank, nranks=world)[0]
    ref = O.ransac(pts, O.EST_HOMOGRAPHY, rng=O.RNG_PHILOX, threshold=2.0, confidence=0.95, max_iterations=4000, seed=3)
    for k in ("inliers", "iterations", "best_hyp"):
        assert r[k] == ref[k], (rank, k, r[k], ref[k])
    assert np.array_equal(r["model"].view(np.uint32), np.asarray(ref["model"], np.float32).view(np.uint32))
    total = D.reduce_sum([float(r["useful_evals"])])[0]
    assert int(total) == ref["evals"], (total, ref["evals"])
    D.barrier()
    ctx.close()
    if rank == 0:
        print(f"MULTI_GPU_OK world={world} inliers={r['inliers']} iterations={r['iterations']}")
    D.finalize()


if __name__ == "__main__":
    main()
