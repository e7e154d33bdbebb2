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
