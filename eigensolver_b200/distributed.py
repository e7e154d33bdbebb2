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
