"""Host-side plumbing of the multi-GPU paths (one process per GPU, torch.distributed): rendezvous, problem sharding,
the hypothesis-sharding index maps, the exchange of the NCCL unique id for the library's own communicator, and the
max/sum reductions bench.py reports. Backend: nccl on a GPU box, gloo on CPU (tests)."""
import os

import numpy as np
import torch
import torch.distributed as dist


def init(backend=None):
    """Join the process group described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun). Returns (rank, world)."""
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world


def bind_to_gpu_numa(local_rank):
    """Pin this process to the CPUs next to its GPU (sysfs local_cpulist of the GPU's PCI function) BEFORE it allocates pinned
    host memory, so that the staging buffers of the end-to-end path live on the GPU's own NUMA node instead of wherever the
    launcher happened to start the rank (first-touch placement). Returns the cpulist string, or None when nothing was done."""
    try:
        p = torch.cuda.get_device_properties(local_rank)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as fh:
            text = fh.read().strip()
        cpus = set()
        for part in text.split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return text
    except Exception:   # noqa: BLE001 - no sysfs entry / not permitted: leave the placement to the launcher
        return None


def _dev():
    return torch.device("cuda", torch.cuda.current_device()) if dist.is_initialized() and dist.get_backend() == "nccl" else torch.device("cpu")


def shard_range(n_items, rank, world):
    """Contiguous block partition of independent problems: rank r owns [lo, hi). Sizes differ by at most one."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def hypothesis_owner(j, world):
    """Hypothesis sharding of one round (usac_fit_cfg.rank/nranks): sample j of the round is solved and scored by rank j % R."""
    return j % world


def gathered_index(j, world, per_rank):
    """Position of sample j's packed score in the all-gathered array [rank][per_rank] (select_kernel's `load`)."""
    return (j % world) * per_rank + j // world


def pack_local_scores(scores_by_sample, rank, world):
    """The slice a rank contributes to the per-round exchange: its samples j = rank, rank + R, ... in order.
    scores_by_sample: uint64/uint2-like array [K, ...] indexed by the sample id within the round (K % R == 0)."""
    return np.ascontiguousarray(scores_by_sample[rank::world])


def unpack_gathered(gathered, world):
    """Inverse of the exchange layout: [R, per_rank, ...] -> [K, ...] in sample order."""
    g = np.asarray(gathered)
    per_rank = g.shape[1]
    out = np.empty((world * per_rank,) + g.shape[2:], g.dtype)
    for r in range(world):
        out[r::world] = g[r]
    return out


def allgather_array(a):
    """All-gather equal-shaped numpy arrays -> [R, ...] (the host emulation of the per-round score exchange)."""
    if not dist.is_initialized():
        return np.asarray(a)[None]
    t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(_dev())
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return np.stack([o.cpu().numpy().view(a.dtype).reshape(a.shape) for o in out])


def broadcast_bytes(payload, src=0):
    """Broadcast a fixed-size byte string (the 128-byte ncclUniqueId of usac_gpu_nccl_unique_id) from `src`."""
    if not dist.is_initialized():
        return payload
    n = 128 if payload is None else len(payload)
    t = torch.zeros(n, dtype=torch.uint8) if payload is None else torch.tensor(list(payload), dtype=torch.uint8)
    t = t.to(_dev())
    dist.broadcast(t, src)
    return bytes(t.cpu().tolist())


def allgather_bytes(payload):
    """All-gather one fixed-size byte string per rank -> list in rank order (the IPC handles of usac_gpu_peer_export)."""
    if not dist.is_initialized():
        return [payload]
    t = torch.tensor(list(payload), dtype=torch.uint8).to(_dev())
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [bytes(o.cpu().tolist()) for o in out]


def reduce_max(values):
    t = torch.tensor(values, dtype=torch.float64, device=_dev())
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def reduce_sum(values):
    t = torch.tensor(values, dtype=torch.float64, device=_dev())
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.tolist()


def barrier():
    if dist.is_initialized():
        dist.barrier()


def finalize():
    if dist.is_initialized():
        dist.destroy_process_group()
