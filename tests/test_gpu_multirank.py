"""Hypothesis sharding (usac_fit_cfg.rank / nranks): results must be identical on every rank and identical to one GPU.
(1) two ranks emulated on ONE GPU: two host threads, two contexts, the per-round exchange done by a host hook installed with
    usac_gpu_set_allgather (device->host, thread barrier, host->device);
(2) real NCCL over NVLink, one process per GPU under torchrun (skipped with fewer than 2 GPUs)."""
import ctypes as C
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

from oracle import oracle as O
from ransac_b200 import generator as gen

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cudart():
    for name in ("libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so.12", "libcudart.so"):
        try:
            return C.CDLL(name)
        except OSError:
            continue
    pytest.skip("libcudart not found")


class HostExchange:
    """usac_allgather_fn for R ranks living in R threads of this process."""
    def __init__(self, world):
        from ransac_b200 import capi
        self.world, self.rt = world, _cudart()
        self.rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        self.rt.cudaStreamSynchronize.argtypes = [C.c_void_p]
        self.barrier = threading.Barrier(world)
        self.parts = [None] * world
        self.calls = 0
        self.fns = [capi.ALLGATHER_FN(self._make(r)) for r in range(world)]

    def _make(self, rank):
        def hook(user, d_send, d_recv, nbytes, stream):
            assert self.rt.cudaStreamSynchronize(stream) == 0
            buf = (C.c_ubyte * nbytes)()
            assert self.rt.cudaMemcpy(buf, d_send, nbytes, 2) == 0            # device -> host
            self.parts[rank] = bytes(buf)
            self.barrier.wait()
            allb = b"".join(self.parts)
            assert self.rt.cudaMemcpy(d_recv, allb, len(allb), 1) == 0        # host -> device
            self.barrier.wait()
            if rank == 0:
                self.calls += 1
            return 0
        return hook


@pytest.mark.parametrize("cfg,world", [(2, 2), (2, 4), (3, 2)])
def test_emulated_ranks_match_single_gpu_and_oracle(cfg, world):
    from ransac_b200 import GpuContext
    est = {2: O.EST_HOMOGRAPHY, 3: O.EST_FUNDAMENTAL}[cfg]
    pts = gen.make(cfg, n=3000)[0]
    thr, conf = gen.CONFIGS[cfg]["threshold"], gen.CONFIGS[cfg]["confidence"]
    ex = HostExchange(world)
    results, errors = [None] * world, []

    def run(rank):
        try:
            ctx = GpuContext(0)
            ctx.set_points(est, pts)
            ctx._check(ctx.L.usac_gpu_set_allgather(ctx.h, ex.fns[rank], None), "set_allgather")
            results[rank] = ctx.fit(thr, conf, 3000, seed=6, round_size=128, rank=rank, nranks=world)[0]
            ctx.close()
        except Exception as e:   # noqa: BLE001
            errors.append(e)
            ex.barrier.abort()
    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(120)
    assert not errors, errors
    assert ex.calls >= 1
    ref = O.ransac(pts, est, rng=O.RNG_PHILOX, threshold=thr, confidence=conf, max_iterations=3000, seed=6)
    for r in results:
        for k in ("inliers", "iterations", "best_hyp", "best_model_idx"):
            assert r[k] == ref[k] == results[0][k], (k, r[k], ref[k])
        assert np.array_equal(r["model"].view(np.uint32), np.asarray(ref["model"], np.float32).view(np.uint32))
    assert sum(r["useful_evals"] for r in results) == ref["evals"]           # the shards partition the sequential work


@pytest.mark.parametrize("cfg,world,batch", [(2, 2, 1), (2, 4, 3), (3, 2, 1)])
def test_peer_windows_match_single_gpu_and_oracle(cfg, world, batch):
    """The exchange over peer memory (usac_gpu_peer_*): ranks = threads of this process, every context owns a window, the reduce
    kernels store into all windows and select_kernel waits for the flags. Same results as one GPU and the oracle, on every rank,
    fit after fit (the sequence numbers keep counting), also for a batch of problems."""
    from ransac_b200 import GpuContext
    est = {2: O.EST_HOMOGRAPHY, 3: O.EST_FUNDAMENTAL}[cfg]
    sets = [gen.make(cfg, seed_offset=40 + b, n=3000 + 500 * b)[0] for b in range(batch)]
    thr, conf = gen.CONFIGS[cfg]["threshold"], gen.CONFIGS[cfg]["confidence"]
    ctxs = [GpuContext(0) for _ in range(world)]
    wins = [c.peer_window() for c in ctxs]
    for r, c in enumerate(ctxs):
        c.set_points(est, np.concatenate(sets), [len(x) for x in sets])
        c.peer_attach_ptrs(wins, r, world)
    results, errors = [[None, None] for _ in range(world)], []

    def run(rank):
        try:
            for rep in range(2):
                results[rank][rep] = ctxs[rank].fit(thr, conf, 3000, seed=6 + rep, round_size=128, rank=rank, nranks=world)
        except Exception as e:   # noqa: BLE001
            errors.append(e)
    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(120)
    for c in ctxs:
        c.close()
    assert not errors, errors
    for rep in range(2):
        for b in range(batch):
            ref = O.ransac(sets[b], est, rng=O.RNG_PHILOX, threshold=thr, confidence=conf, max_iterations=3000, seed=6 + rep)
            for rank in range(world):
                r = results[rank][rep][b]
                for k in ("inliers", "iterations", "best_hyp", "best_model_idx"):
                    assert r[k] == ref[k], (rep, b, rank, k, r[k], ref[k])
                assert np.array_equal(r["model"].view(np.uint32), np.asarray(ref["model"], np.float32).view(np.uint32))
            assert sum(results[rank][rep][b]["useful_evals"] for rank in range(world)) == ref["evals"]


@pytest.mark.parametrize("sparse", [False, True])
def test_peer_windows_napsac_grid(sparse):
    """NAPSAC over a grid with the hypotheses sharded over 3 ranks (peer windows): every rank computes the seeds / cursors of ALL samples
    and solves its own; `sparse` = data with one usable cell, where the sampler switches to uniform sampling part-way (do_uniform)."""
    from ransac_b200 import GpuContext
    from ransac_b200.api import NEIGH_GRID, SAMPLER_NAPSAC
    world = 3
    if sparse:
        g = np.random.default_rng(78)
        pts = (g.random((900, 4)) * 1000).astype(np.float32)
        pts[:6] = np.float32([310.0, 420.0, 615.0, 120.0]) + (g.random((6, 4)) * 5).astype(np.float32)
    else:
        pts = gen.homography(n=20000, inlier_ratio=0.1, clustered=True, seed=31)[0]
    ctxs = [GpuContext(0) for _ in range(world)]
    wins = [c.peer_window() for c in ctxs]
    for r, c in enumerate(ctxs):
        c.set_points(O.EST_HOMOGRAPHY, pts)
        c.set_neighbors_grid(0, 50)
        c.peer_attach_ptrs(wins, r, world)
    results, errors = [None] * world, []

    def run(rank):
        try:
            results[rank] = ctxs[rank].fit(2.0, 0.95, 1500, sampler=SAMPLER_NAPSAC, neighbors=NEIGH_GRID, seed=4, round_size=129, rank=rank, nranks=world)[0]
        except Exception as e:   # noqa: BLE001
            errors.append(e)
    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(120)
    for c in ctxs:
        c.close()
    assert not errors, errors
    ref = O.ransac(pts, O.EST_HOMOGRAPHY, sampler=O.SAMPLER_NAPSAC, rng=O.RNG_PHILOX, neighbors=O.NEIGH_GRID, cell_size=50,
                   threshold=2.0, confidence=0.95, max_iterations=1500, seed=4)
    for r in results:
        for k in ("inliers", "iterations", "best_hyp", "best_model_idx"):
            assert r[k] == ref[k], (k, r[k], ref[k])
        assert np.array_equal(r["model"].view(np.uint32), np.asarray(ref["model"], np.float32).view(np.uint32))


def test_config5_emulated_ranks_full_size():
    """BASELINE config 5 at its stated size (1M correspondences, NAPSAC grid): the hypotheses of every round split over 4
    emulated ranks give the single-GPU result on every rank (which test_gpu_parity checks against the oracle)."""
    from ransac_b200 import GpuContext
    from ransac_b200.api import NEIGH_GRID, SAMPLER_NAPSAC
    pts = gen.make(5)[0]
    world = 4
    kw = dict(sampler=SAMPLER_NAPSAC, neighbors=NEIGH_GRID, seed=1, round_size=1024)
    one = GpuContext(0)
    one.set_points(O.EST_HOMOGRAPHY, pts)
    one.set_neighbors_grid(0, 50)
    single = one.fit(2.0, 0.95, 4096, **kw)[0]
    one.close()
    ex = HostExchange(world)
    results, errors = [None] * world, []

    def run(rank):
        try:
            ctx = GpuContext(0)
            ctx.set_points(O.EST_HOMOGRAPHY, pts)
            ctx.set_neighbors_grid(0, 50)
            ctx._check(ctx.L.usac_gpu_set_allgather(ctx.h, ex.fns[rank], None), "set_allgather")
            results[rank] = ctx.fit(2.0, 0.95, 4096, rank=rank, nranks=world, **kw)[0]
            ctx.close()
        except Exception as e:   # noqa: BLE001
            errors.append(e)
            ex.barrier.abort()
    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(300)
    assert not errors, errors
    for r in results:
        for k in ("inliers", "iterations", "best_hyp", "best_model_idx"):
            assert r[k] == single[k], (k, r[k], single[k])
        assert np.array_equal(r["model"].view(np.uint32), single["model"].view(np.uint32))
    assert sum(r["useful_evals"] for r in results) == single["useful_evals"]


def test_nccl_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "MULTI_GPU_OK" in r.stdout, r.stdout[-3000:]
