"""world_size-2 gloo tests (CPU) of the host-side logic of the N > 1 paths: rendezvous, problem sharding, the
hypothesis-sharding exchange layout and the reductions. The per-sample scores come from the CPU oracle here; the
select step is restated in Python with the sequential semantics of ransac.cpp:58-139 (prefix best + first stop)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _select(scores, n, m, conf, max_it):
    """scores: [K, 2] (inliers, sum) in sample order, -1 = no model. -> (best sample, inliers, iterations)."""
    from oracle import oracle as O
    best, best_key, iters, max_iters = -1, (0, 0.0), 0, max_it
    for j, (c, s) in enumerate(scores):
        if not iters < max_iters:
            break
        if c >= 0 and (c, s) > best_key:
            best, best_key = j, (c, s)
            max_iters = O.standard_termination(int(c), n, m, conf, max_it)
        iters += 1
    return best, int(best_key[0]), iters


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from oracle import oracle as O
    from ransac_b200 import dist as D
    from ransac_b200 import generator as gen
    r, w = D.init("gloo")
    assert (r, w) == (rank, world)
    # --- problem sharding + reductions -------------------------------------------------------------------------
    lo, hi = D.shard_range(7, rank, world)
    assert D.reduce_sum([hi - lo])[0] == 7
    assert D.reduce_max([10.0 + rank])[0] == 10.0 + world - 1
    uid = D.broadcast_bytes(bytes(range(128)) if rank == 0 else None)
    assert uid == bytes(range(128))
    handles = D.allgather_bytes(bytes([rank + 1]) * 64)              # the exchange of the peer-window IPC handles (usac_gpu_peer_export)
    assert handles == [bytes([r + 1]) * 64 for r in range(world)]
    # --- hypothesis sharding of one round: each rank scores samples j % R == rank ---------------------------------
    pts, _, _ = gen.homography(n=600, seed=21)
    K, n, m = 64, len(pts), 4
    samples = np.stack([O.philox_unique(9, j, 0, n, m) for j in range(K)])
    local = np.full((K, 2), -1.0)
    for j in range(K):
        if D.hypothesis_owner(j, world) != rank:
            continue
        mods = O.solve_minimal(O.EST_HOMOGRAPHY, pts, samples[j])
        if len(mods):
            c, s, _ = O.score(O.EST_HOMOGRAPHY, pts, mods[0], 2.0)
            local[j] = (c, s)
    mine = D.pack_local_scores(local, rank, world)
    assert mine.shape == (K // world, 2)
    gathered = D.allgather_array(mine)
    full = D.unpack_gathered(gathered, world)
    for j in range(K):
        assert np.array_equal(gathered.reshape(-1, 2)[D.gathered_index(j, world, K // world)], full[j])
    best, inl, iters = _select(full, n, m, 0.95, K)
    ref = O.ransac(pts, O.EST_HOMOGRAPHY, rng=O.RNG_PHILOX, threshold=2.0, confidence=0.95, max_iterations=K, seed=9)
    assert (best, inl, iters) == (ref["best_hyp"], ref["inliers"], ref["iterations"]), ((best, inl, iters), ref)
    out[rank] = (best, inl, iters)
    D.barrier()
    D.finalize()


def test_two_rank_gloo_host_logic():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert out[0] == out[1]


def test_shard_range_covers_everything():
    from ransac_b200 import dist as D
    for n in (0, 1, 7, 8, 2368):
        for w in (1, 2, 3, 8):
            ranges = [D.shard_range(n, r, w) for r in range(w)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            assert max(h - l for l, h in ranges) - min(h - l for l, h in ranges) <= 1


def test_bind_to_gpu_numa_is_harmless_without_a_gpu():
    import os

    from ransac_b200 import dist as D
    before = os.sched_getaffinity(0)
    assert D.bind_to_gpu_numa(0) is None or isinstance(D.bind_to_gpu_numa(0), str)
    assert os.sched_getaffinity(0) <= before
