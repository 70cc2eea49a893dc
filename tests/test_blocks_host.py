"""Host side of the block-sharded entry path (BASELINE config C5) on CPU: the LZ4 frame index walk
(zpb_lz4_frame_index — pure framing, no GPU), the block partition, and a world_size-2 gloo run of the
64-byte accumulator relay."""
import os
import socket

import numpy as np
import pytest

from zpack_b200 import corpus, shard
from zpack_b200 import lib as zlib


def _walk(frame):
    """independent restatement of the block-header chain for the test (lz4_Frame_format.md)"""
    assert frame[:4].tobytes() == b"\x04\x22\x4d\x18"
    flg = int(frame[4])
    p = 7 + (8 if flg & 8 else 0) + (4 if flg & 1 else 0)
    out = []
    while True:
        bh = int.from_bytes(frame[p:p + 4].tobytes(), "little")
        p += 4
        if bh == 0:
            break
        out.append((p, bh & 0x7FFFFFFF, bh >> 31))
        p += bh & 0x7FFFFFFF
    assert p == len(frame)
    return out


def test_frame_index_matches_the_block_header_chain(oracle):
    data = corpus.big_entry(5 * 65536 + 1234, piece=2 * 65536)      # random | text | runs pieces, short last block
    frame = oracle.lz4f_encode_port(data, 0, independent=True)
    got = zlib.lz4_frame_index(frame, archive_off=10)
    assert got is not None
    blocks, bs, content = got
    want = _walk(frame)
    assert bs == 65536 and content is None and len(blocks) == len(want) == 6
    assert [(int(b["src_off"]) - 10, int(b["comp_size"]), int(b["flags"])) for b in blocks] == want
    assert blocks["flags"][0] == zlib.BLK_STORED                   # the random piece is stored


def test_frame_index_declines_what_the_sharded_path_does_not_take(oracle, lz4_cases):
    data = corpus.big_entry(3 * 65536, piece=65536, first=1)
    linked = oracle.lz4f_encode_port(data, 0, independent=False)    # the reference writer's shape (FLG 0x40)
    assert zlib.lz4_frame_index(linked) is None
    ok = oracle.lz4f_encode_port(data, 0, independent=True)
    assert zlib.lz4_frame_index(ok) is not None
    for cut in (3, 10, len(ok) - 1, len(ok) - 5):
        assert zlib.lz4_frame_index(ok[:cut]) is None              # truncated: no EndMark where the chain ends
    bad = ok.copy()
    bad[6] ^= 0x55                                                  # header checksum byte (lz4frame.c:1184-1186)
    assert zlib.lz4_frame_index(bad) is None
    assert zlib.lz4_frame_index(np.concatenate([ok, ok])) is None   # two frames in one entry: general path
    # every reference-written fixture: accepted iff independent, 64 KB blocks, no checksums
    seen = 0
    for name, comp in lz4_cases.items():
        if name.endswith("__in") or len(comp) < 11 or comp[:4].tobytes() != b"\x04\x22\x4d\x18":
            continue
        seen += 1
        flg, bd = int(comp[4]), int(comp[5])
        expect = bool(flg & 0x20) and not (flg & 0x14) and ((bd >> 4) & 7) == 4 and _single_frame(comp)
        got = zlib.lz4_frame_index(comp)
        assert (got is not None) == expect, name
        if got is not None:
            assert [(int(b["src_off"]), int(b["comp_size"]), int(b["flags"])) for b in got[0]] == _walk(comp)
    assert seen >= 20


def _single_frame(comp):
    try:
        _walk(comp)
        return True
    except (AssertionError, IndexError, ValueError):
        return False


def test_split_blocks():
    for n, world in [(32768, 8), (262144, 8), (7, 2), (1, 4), (9, 8), (3, 8)]:
        total = n * 65536 - 65536 + 100                            # short last block
        parts = shard.split_blocks(n, world, total_size=total)
        assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == n
        for (a, b), (c, d) in zip(parts, parts[1:]):
            assert a <= b == c <= d
        last = [p for p in parts if p[1] > p[0]][-1]
        assert last[1] == n and (n == 1 or last[1] - last[0] >= 2)
    assert shard.split_blocks(262144, 8) == [(k * 32768, (k + 1) * 32768) for k in range(8)]   # C5: 2 GiB per GPU


def _relay_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from zpack_b200 import shard as S
    parts = S.split_blocks(5, world, total_size=4 * 65536 + 10)

    def send(dst, acc):
        t = torch.zeros(9, dtype=torch.int64)
        if acc is not None:
            t[0] = 1
            t[1:] = torch.from_numpy(np.asarray(acc, np.uint64).view(np.int64))
        dist.send(t, dst)

    def recv(src):
        t = torch.zeros(9, dtype=torch.int64)
        dist.recv(t, src)
        return t[1:].numpy().view(np.uint64).copy() if int(t[0]) else None

    def chain(acc):   # stand-in for zpb_blocks_digest: order-sensitive, so a wrong relay order shows
        a = np.arange(8, dtype=np.uint64) if acc is None else acc
        a = a * np.uint64(0x9E3779B1) + np.uint64(parts[rank][0] + 1)
        return a, (int(a.sum() & np.uint64(0xFFFFFFFF)) if parts[rank][1] == 5 else None)

    dg = S.relay_digest(rank, world, chain, send, recv, empty=parts[rank][1] == parts[rank][0])
    dist.barrier()
    q.put((rank, dg))
    dist.destroy_process_group()


def test_two_rank_gloo_accumulator_relay():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_relay_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = dict(q.get(timeout=120) for _ in ps)
    [p.join(60) for p in ps]
    parts = shard.split_blocks(5, 2, total_size=4 * 65536 + 10)
    a = np.arange(8, dtype=np.uint64)
    for lo, hi in parts:
        a = a * np.uint64(0x9E3779B1) + np.uint64(lo + 1)
    assert res[0] is None and res[1] == int(a.sum() & np.uint64(0xFFFFFFFF))
