"""Host-side multi-GPU logic on CPU: the entry partition, and a world_size-2 gloo run of the same
gather / max-over-ranks plumbing bench.py uses under torchrun (no data-path collective exists)."""
import os
import socket
import sys

import numpy as np
import pytest

from zpack_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_is_contiguous_covering_and_balanced():
    rng = np.random.default_rng(1)
    for n, world in [(0, 4), (1, 8), (7, 8), (1000, 2), (65536, 8), (12345, 3)]:
        u = rng.integers(0, 200000, size=n)
        parts = shard.partition(u, world)
        assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == n
        for (a, b), (c, d) in zip(parts, parts[1:]):
            assert a <= b == c <= d
        if n >= 50 * world:
            loads = [u[a:b].sum() for a, b in parts]
            assert max(loads) <= 1.05 * (u.sum() / world) + 200000
    same = shard.partition(np.full(65536, 131072), 8)
    assert [b - a for a, b in same] == [8192] * 8                     # C2: 8192 entries, 1 GiB per GPU


def test_byte_range_and_pack_offsets():
    off = np.array([10, 110, 410], np.uint64)
    cs = np.array([100, 300, 50], np.uint64)
    assert shard.byte_range(off, cs, 1, 3) == (110, 460)
    assert shard.byte_range(off, cs, 2, 2) == (0, 0)
    assert list(shard.pack_offsets(cs)) == [10, 110, 410]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    u = np.full(1000, 131072)
    a, b = shard.partition(u, world)[rank]
    # each rank "unpacks" its shard: status/digest stay local; only timing and counts are reduced
    t = torch.tensor([float(10 + rank), float(b - a)], dtype=torch.float64)
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = t.clone()
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    dist.barrier()
    q.put((rank, a, b, float(tmax[0]), float(tsum[1])))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(60) for p in ps]
    assert [(r[1], r[2]) for r in res] == [(0, 500), (500, 1000)]
    assert all(r[3] == 11.0 and r[4] == 1000.0 for r in res)          # max-over-ranks time, whole-job count


def test_c_abi_partition_equals_the_python_rule():
    """zpb_group_partition (C++, what libzpack.so's batched read uses) against shard.partition on the same sizes:
    contiguous, covering, and balanced to within one entry of the ideal split."""
    from zpack_b200 import lib as zlib
    rng = np.random.default_rng(7)
    for n, world in ((0, 3), (1, 4), (17, 2), (1000, 8), (65536, 8), (5, 8)):
        e = np.zeros(n, zlib.Entry)
        e["uncomp_size"] = rng.integers(0, 300000, n)
        e["comp_size"] = np.maximum(1, e["uncomp_size"] // 2)
        e["src_off"] = 10 + np.concatenate([[0], np.cumsum(e["comp_size"])[:-1]]) if n else 0
        perm = rng.permutation(n)                       # callers need not pass archive order
        order, cuts = zlib.group_partition(np.ascontiguousarray(e[perm]), world)
        assert cuts[0] == 0 and cuts[-1] == n and (np.diff(cuts.astype(np.int64)) >= 0).all()
        assert sorted(order.tolist()) == list(range(n))
        assert (np.diff(e[perm]["src_off"][order.astype(np.int64)].astype(np.int64)) >= 0).all()   # archive order
        if n >= world * 4:
            total = float(e["uncomp_size"].sum())
            sizes = e[perm]["uncomp_size"][order.astype(np.int64)]
            for k in range(world):
                share = float(sizes[int(cuts[k]):int(cuts[k + 1])].sum())
                assert abs(share - total / world) <= 300000 + 1, (n, world, k, share, total / world)
