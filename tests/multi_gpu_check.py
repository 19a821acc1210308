"""torchrun entry (one process per GPU): hypothesis-sharded fit of one large homography problem, first with one NCCL all-gather per round,
then with the exchange over peer memory (CUDA IPC windows, usac_gpu_peer_*); every rank must produce the oracle's result both ways.
Launched by tests/test_gpu_multirank.py::test_nccl_two_gpus and usable by hand:
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
    ref = O.ransac(pts, O.EST_HOMOGRAPHY, rng=O.RNG_PHILOX, threshold=2.0, confidence=0.95, max_iterations=4000, seed=3)

    def check(r, what):
        for k in ("inliers", "iterations", "best_hyp"):
            assert r[k] == ref[k], (what, rank, k, r[k], ref[k])
        assert np.array_equal(r["model"].view(np.uint32), np.asarray(ref["model"], np.float32).view(np.uint32)), what
        total = D.reduce_sum([float(r["useful_evals"])])[0]
        assert int(total) == ref["evals"], (what, total, ref["evals"])

    # (1) NCCL all-gather per round
    ctx.nccl_init(uid, rank, world)
    check(ctx.fit(2.0, 0.95, 4000, seed=3, round_size=512, rank=rank, nranks=world)[0], "nccl")
    D.barrier()
    # (2) peer windows: attached windows take precedence over the hook; two fits (the sequence numbers keep counting)
    ctx.peer_attach(D.allgather_bytes(ctx.peer_export()), rank, world)
    D.barrier()                                                       # nobody waits for a rank that is still mapping windows
    for rep in range(2):
        check(ctx.fit(2.0, 0.95, 4000, seed=3, round_size=512, rank=rank, nranks=world)[0], f"peer {rep}")
    D.barrier()
    ctx.close()
    if rank == 0:
        print(f"MULTI_GPU_OK world={world} inliers={ref['inliers']} iterations={ref['iterations']} exchange=nccl+peer")
    D.finalize()


if __name__ == "__main__":
    main()
