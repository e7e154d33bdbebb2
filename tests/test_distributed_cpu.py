"""Host-side multi-rank logic on CPU: world_size 2, gloo.  Each rank sweeps its k shard
(here with the C oracle standing in for the GPU, tests only) and the root tables are
gathered with eigensolver_b200.distributed.gather_root_tables; the result must equal the
single-process table."""
import os
import socket

import numpy as np
import pytest

from eigensolver_b200.distributed import gather_root_tables, shard_bounds, shard_k
from oracle import rk_oracle as ork


def test_shard_bounds_cover_exactly():
    for n in (1, 7, 8, 1000, 1001):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _local_table(k, W, offset_unused=0):
    model = ork.make_model("cylinder_density", n_ext=1500, n_int=96)
    e, i = ork.grid(model, 1, k, W, threads=2)
    ki, wi = ork.brackets(e - i)
    omega = np.array([0.5 * (k[a] * W[b] + k[a] * W[b + 1]) for a, b in zip(ki, wi)])
    acc = (np.arange(len(ki)) % 2).astype(np.int32)
    return ki, omega, acc


def _worker(rank, world, port, k, W, out, layout="contiguous"):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ks, off, stride = shard_k(k, rank, world, layout)
        ki, om, acc = _local_table(ks, W)
        # the accepted flag of this test is a function of the GLOBAL row so that it survives re-sharding
        acc = ((ki * stride + off) % 2).astype(np.int32)
        gk, gw, ga = gather_root_tables(ki, om, acc, off, k_stride=stride)
        if rank == 0:
            np.savez(out, k=gk, w=gw, a=ga)
        else:
            # every rank receives the same full table
            assert len(gk) == len(gw) == len(ga)
    finally:
        dist.destroy_process_group()


def test_strided_shards_cover_exactly():
    k = np.arange(11.0)
    seen = []
    for r in range(3):
        ks, off, stride = shard_k(k, r, 3, "strided")
        assert (off, stride) == (r, 3)
        seen += list(off + stride * np.arange(len(ks)))
        assert np.array_equal(ks, k[off + stride * np.arange(len(ks))])
    assert sorted(seen) == list(range(11))


@pytest.mark.timeout(300)
@pytest.mark.parametrize("layout", ["contiguous", "strided"])
def test_two_rank_gloo_gather_equals_single_process(tmp_path, layout):
    import torch.multiprocessing as mp
    k = np.linspace(0.5, 4.0, 9)          # odd: uneven shards
    W = np.linspace(2.95, 4.95, 48)
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, _free_port(), k, W, out, layout), nprocs=2, join=True)
    g = np.load(out)
    ki, om, acc = _local_table(k, W)
    acc = (ki % 2).astype(np.int32)
    assert len(ki) > 0
    assert np.array_equal(g["k"], ki)
    assert np.array_equal(g["w"], om)
    assert np.array_equal(g["a"], acc)


def test_gather_handles_empty_rank(tmp_path):
    """A rank whose shard has no brackets contributes an empty table."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(_free_port())
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        gk, gw, ga = gather_root_tables(np.zeros(0, np.int32), np.zeros(0), np.zeros(0, np.int32), 5)
        assert len(gk) == len(gw) == len(ga) == 0
        gk, gw, ga = gather_root_tables(np.array([0, 2], np.int32), np.array([1.5, 2.5]),
                                        np.array([1, 0], np.int32), 5)
        assert list(gk) == [5, 7] and list(gw) == [1.5, 2.5] and list(ga) == [1, 0]
    finally:
        dist.destroy_process_group()


def _scan_worker(rank, world, port, out):
    """two ranks, each with the compact scan table of its strided wavenumbers (synthetic tables here:
    the gather is host-side logic)"""
    import torch.distributed as dist
    from eigensolver_b200.scan import ScanResult, gather_scan_modes
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7 + rank)
        n = 5 + 3 * rank                     # uneven shares; rank 1 also has unaccepted entries
        table = {"model": rng.integers(0, 3, n).astype(np.int32), "slot": rng.integers(0, 2, n).astype(np.int32),
                 "k_index": rng.integers(0, 4, n).astype(np.int32), "omega": rng.uniform(1, 2, n),
                 "accepted": (np.arange(n) % (2 if rank else 1) == 0).astype(np.int32)}
        res = ScanResult([], table, np.zeros(4), rank, world)
        g = gather_scan_modes(res, "cpu").numpy()
        mine = table["accepted"] == 1
        want = np.stack([table["model"][mine], table["slot"][mine], table["k_index"][mine] * world + rank,
                         table["omega"][mine]], axis=1)
        np.savez(out % rank, gathered=g, mine=want)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gather_of_scan_modes(tmp_path):
    """gather_scan_modes: the accepted modes of every rank, global rows = k_offset + k_index * k_stride,
    concatenated in rank order; every rank receives the same table."""
    import torch.multiprocessing as mp
    out = str(tmp_path / "scan%d.npz")
    mp.spawn(_scan_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    g0, g1 = np.load(out % 0), np.load(out % 1)
    assert np.array_equal(g0["gathered"], g1["gathered"])
    assert np.array_equal(g0["gathered"], np.concatenate([g0["mine"], g1["mine"]]))
    assert len(g1["mine"]) < 8            # the unaccepted entries of rank 1 were left out
