"""Multi-GPU plumbing: the k axis shards across ranks, the root tables are gathered.

The reference parallelises by starting one OS process per (k, speed interval)
(Density_cylinder.py:1142-1159) and collecting the per-process lists through
multiprocessing queues (:1161-1171).  Here the same decomposition is one process
per GPU: every rank sweeps its own contiguous slab of wavenumbers - no data-path
collective, the (k, omega) points are independent - and the only exchange is the
gather of the (small) root tables, done with torch.distributed (NCCL over NVLink on
GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n, rank, world):
    """Contiguous, balanced [lo, hi) slice of n items for `rank` of `world`."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_k(k, rank, world):
    lo, hi = shard_bounds(len(k), rank, world)
    return np.asarray(k)[lo:hi], lo


def gather_root_tables(k_index, omega, accepted, k_offset, device=None, group=None):
    """All-gather variable-length root tables.

    k_index is local to the rank's shard; `k_offset` (the shard's first global row) makes
    it global.  Returns (k_index, omega, accepted) of ALL ranks, ordered by rank, i.e.
    sorted by global k index.  Every rank receives the full table."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    dev = device if device is not None else "cpu"
    n = torch.tensor([len(omega)], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    # one packed fp64 payload per rank: [global k index, omega, accepted]
    pay = torch.zeros((cap, 3), dtype=torch.float64, device=dev)
    if len(omega):
        pay[: len(omega), 0] = torch.as_tensor(np.asarray(k_index, dtype=np.float64) + k_offset, device=dev)
        pay[: len(omega), 1] = torch.as_tensor(np.asarray(omega, dtype=np.float64), device=dev)
        pay[: len(omega), 2] = torch.as_tensor(np.asarray(accepted, dtype=np.float64), device=dev)
    bufs = [torch.zeros_like(pay) for _ in range(world)]
    dist.all_gather(bufs, pay, group=group)
    parts = [b[:c].cpu().numpy() for b, c in zip(bufs, counts)]
    allp = np.concatenate(parts, axis=0) if parts else np.zeros((0, 3))
    return allp[:, 0].astype(np.int64), allp[:, 1].copy(), allp[:, 2].astype(np.int32)


class _DevArray:
    """Zero-copy view of library-owned device memory for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def gather_root_tables_device(solver, slot, k_offset, device, group=None):
    """All-gather the root table of mode slot `slot` straight from the solver's device buffers
    (no host round trip): NCCL over NVLink moves (global k index, omega, accepted) of every rank.
    Returns a float64 tensor [total, 3] on `device`, ordered by rank = sorted by global k index."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    info = solver.roots_device(slot)
    n = info["n"]
    cnt = torch.tensor([n], dtype=torch.int64, device=device)
    counts = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(counts, cnt, group=group)
    counts = counts.tolist()
    cap = max(max(counts), 1)
    pay = torch.zeros((cap, 3), dtype=torch.float64, device=device)
    if n:
        ki = torch.as_tensor(_DevArray(info["k_index"][0], n, "<i4"), device=device)
        om = torch.as_tensor(_DevArray(info["omega"][0], n, "<f8"), device=device)
        ac = torch.as_tensor(_DevArray(info["accepted"][0], n, "<i4"), device=device)
        pay[:n, 0] = ki.to(torch.float64) + float(k_offset)
        pay[:n, 1] = om
        pay[:n, 2] = ac.to(torch.float64)
    out = torch.empty((world * cap, 3), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out, pay, group=group)
    parts = [out[r * cap: r * cap + c] for r, c in enumerate(counts)]
    return torch.cat(parts, dim=0)
