"""Multi-GPU plumbing: the k axis shards across ranks, the root tables are gathered.

The reference parallelises by starting one OS process per (k, speed interval)
(Density_cylinder.py:1142-1159) and collecting the per-process lists through
multiprocessing queues (:1161-1171).  Here the same decomposition is one process
per GPU: every rank sweeps its own share of the wavenumbers - no data-path
collective, the (k, omega) points are independent - and the only exchange is the
gather of the (small) root tables, done with torch.distributed (NCCL over NVLink on
GPUs, gloo in the CPU tests).

Sharding: "strided" (rank r owns k[r::world]) is the default of the bench - the number of
modes grows with k, so contiguous slabs give the last rank ~2x the brackets of the first and
the step time is the maximum over ranks; "contiguous" slabs are kept for callers that want
them.  Either way a local row i of rank r is the global row  k_offset + i * k_stride.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n, rank, world):
    """Contiguous, balanced [lo, hi) slice of n items for `rank` of `world`."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_k(k, rank, world, layout="contiguous"):
    """-> (k of this rank, k_offset, k_stride): local row i is global row k_offset + i*k_stride."""
    k = np.asarray(k)
    if layout == "strided":
        return np.ascontiguousarray(k[rank::world]), rank, world
    lo, hi = shard_bounds(len(k), rank, world)
    return k[lo:hi], lo, 1


def _sorted_by_global_row(allp):
    """Rows of [global k index, omega, ...] ordered by (k index, omega): the order of the
    single-process table (brackets of one k-row are already ascending in omega)."""
    order = np.lexsort((allp[:, 1], allp[:, 0]))
    return allp[order]


def gather_root_tables(k_index, omega, accepted, k_offset, device=None, group=None, k_stride=1):
    """All-gather variable-length root tables (host arrays in, host arrays out).

    k_index is local to the rank's shard; global row = k_offset + k_index * k_stride.
    Returns (k_index, omega, accepted) of ALL ranks sorted by global k index, then omega,
    i.e. the single-process table.  Every rank receives the full table."""
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
        gk = np.asarray(k_index, dtype=np.float64) * k_stride + k_offset
        pay[: len(omega), 0] = torch.as_tensor(gk, device=dev)
        pay[: len(omega), 1] = torch.as_tensor(np.asarray(omega, dtype=np.float64), device=dev)
        pay[: len(omega), 2] = torch.as_tensor(np.asarray(accepted, dtype=np.float64), device=dev)
    bufs = [torch.zeros_like(pay) for _ in range(world)]
    dist.all_gather(bufs, pay, group=group)
    parts = [b[:c].cpu().numpy() for b, c in zip(bufs, counts)]
    allp = np.concatenate(parts, axis=0) if parts else np.zeros((0, 3))
    if k_stride != 1:
        allp = _sorted_by_global_row(allp)
    return allp[:, 0].astype(np.int64), allp[:, 1].copy(), allp[:, 2].astype(np.int32)


def _consumer_stream(device):
    """cudaStream_t of torch's current stream on `device` (None on a CPU device: the call blocks)."""
    import torch
    dev = torch.device(device)
    if dev.type != "cuda":
        return None
    # the legacy default stream has handle 0, which esb_tables_wait reads as "block the host"
    return torch.cuda.current_stream(dev).cuda_stream or None


class _DevArray:
    """Zero-copy view of library-owned device memory for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def gather_root_tables_device(solver, slot, k_offset, device, group=None, k_stride=1, accepted_only=False,
                              sort=False):
    """All-gather the root table of mode slot `slot` straight from the solver's device buffers
    (no host round trip): NCCL over NVLink moves (global k index, omega, accepted) of every rank.

    accepted_only: gather only the modes (what the reference's sol_ks / sol_omegas hold), payload
    [global k index, omega]; otherwise every bracket with its accepted flag, payload
    [global k index, omega, accepted].  sort: order by (global k index, omega) on the device
    (needed for strided shards; contiguous shards are already in that order).
    Returns a float64 tensor [total, 2 or 3] on `device`."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    # torch reads the library's buffers on its current stream: order that stream after the sweep
    info = solver.roots_device(slot, stream=_consumer_stream(device))
    n = info["n"]
    cols = 2 if accepted_only else 3
    if n:
        ki = torch.as_tensor(_DevArray(info["k_index"][0], n, "<i4"), device=device)
        om = torch.as_tensor(_DevArray(info["omega"][0], n, "<f8"), device=device)
        ac = torch.as_tensor(_DevArray(info["accepted"][0], n, "<i4"), device=device)
        gk = ki.to(torch.float64) * float(k_stride) + float(k_offset)
        if accepted_only:
            m = ac == 1
            mine = torch.stack((gk[m], om[m]), dim=1)
        else:
            mine = torch.stack((gk, om, ac.to(torch.float64)), dim=1)
    else:
        mine = torch.zeros((0, cols), dtype=torch.float64, device=device)
    cnt = torch.tensor([mine.shape[0]], dtype=torch.int64, device=device)
    counts = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(counts, cnt, group=group)
    counts = counts.tolist()
    cap = max(max(counts), 1)
    pay = torch.zeros((cap, cols), dtype=torch.float64, device=device)
    pay[: mine.shape[0]] = mine
    out = torch.empty((world * cap, cols), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out, pay, group=group)
    parts = [out[r * cap: r * cap + c] for r, c in enumerate(counts)]
    full = torch.cat(parts, dim=0)
    if sort and full.shape[0]:
        # stable two-key sort: omega first, then the global row
        full = full[torch.argsort(full[:, 1], stable=True)]
        full = full[torch.argsort(full[:, 0], stable=True)]
    return full


def gather_modes_device(solver, n_slots, k_offset, device, group=None, k_stride=1, sort=False):
    """All-gather the accepted modes of ALL mode slots in one exchange: the concatenation of the
    reference's sol_ks / sol_omegas lists of every rank, read straight from the solver's device
    buffers.  Returns a float64 tensor [total, 3] = (global k row, omega, slot) on `device`, ordered
    by rank (like the reference's queue output, which is in process order) or, with sort=True, by
    (slot, global k row, omega).

    The payload is packed by the library (esb_pack_modes_dev: two small launches, deterministic order,
    row 0 = the count) into a buffer of a capacity all ranks share, so a step is ONE fixed-size
    all_gather_into_tensor - no count exchange, no boolean-mask compaction, no host synchronisation before
    the collective is enqueued.  The capacity is negotiated once (all_reduce MAX of the table sizes) and
    again only when some rank reports that it was exceeded."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from . import _lib as L

    world = dist.get_world_size(group)
    stream = _consumer_stream(device)
    st = solver.__dict__.setdefault("_gather_state", {})

    def allocate(cap):
        st.update(cap=cap, n_slots=n_slots, world=world,
                  pay=torch.zeros((cap + 1, 3), dtype=torch.float64, device=device),
                  out=torch.empty((world, cap + 1, 3), dtype=torch.float64, device=device))

    def negotiate(minimum):
        out, n = L.esb_roots(), C.c_int32(0)
        bound = 0
        for slot in range(n_slots):           # table sizes are host knowledge of the sweep call: no wait
            L.check(solver.lib, solver.ctx, solver.lib.esb_roots_device(solver.ctx, slot, C.byref(out), C.byref(n)),
                    "esb_roots_device")
            bound += n.value
        want = torch.tensor([max(bound, minimum)], dtype=torch.int64, device=device)
        dist.all_reduce(want, op=dist.ReduceOp.MAX, group=group)
        allocate(int(want.item() * 1.25) + 1024)

    if st.get("n_slots") != n_slots or st.get("world") != world or st["pay"].device != torch.device(device):
        negotiate(0)
    while True:
        L.check(solver.lib, solver.ctx,
                solver.lib.esb_pack_modes_dev(solver.ctx, int(n_slots), float(k_offset), float(k_stride),
                                              C.c_void_p(st["pay"].data_ptr()), st["cap"],
                                              C.c_void_p(int(stream)) if stream else None),
                "esb_pack_modes_dev")
        # (stream None = torch on the legacy default stream: it orders itself after the context's blocking
        #  stream without an event)
        dist.all_gather_into_tensor(st["out"].view(-1, 3), st["pay"], group=group)
        head = st["out"][:, 0, :].tolist()       # the one host wait of the step, after the collective
        if not any(h[2] for h in head):
            break
        negotiate(int(max(h[1] for h in head)))  # some rank found more modes than the shared capacity
    full = torch.cat([st["out"][r, 1: 1 + int(h[0])] for r, h in enumerate(head)], dim=0)
    # the first capacity comes from the bracket counts, an upper bound ~2x the modes: every rank sees the same
    # headers, so all of them can tighten it to what the modes need (half the bytes of the next exchanges)
    need = int(max(h[0] for h in head) * 1.25) + 1024
    if st["cap"] > 1.5 * need:
        allocate(need)
    if sort and full.shape[0]:
        for col in (1, 0, 2):            # stable sorts, least significant key first
            full = full[torch.argsort(full[:, col], stable=True)]
    return full
