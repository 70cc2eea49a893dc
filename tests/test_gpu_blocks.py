"""GPU parity for the block-sharded entry path (BASELINE config C5): one LZ4 entry with
independent 64 KB blocks, decoded one warp per block through zpb_unpack_blocks_device, its XXH3-64 chained
by zpb_blocks_digest — against the oracle's bytes and digest, in one shard and relayed over several."""
import numpy as np
import pytest

from zpack_b200 import container, corpus, shard
from zpack_b200 import lib as zlib

pytestmark = pytest.mark.gpu


def _decode_sharded(ctx, frame, total, world, archive_off=0, pad=0):
    """world shards, decoded and chained in shard order on one GPU (what N ranks do with a 64-byte relay)."""
    import torch
    idx = zlib.lz4_frame_index(frame, archive_off=archive_off)
    assert idx is not None
    blocks, bs, _ = idx
    arch = np.concatenate([np.zeros(archive_off, np.uint8), frame, np.zeros(pad, np.uint8)])
    d_arch = torch.from_numpy(arch).cuda()
    out = np.zeros(total, np.uint8)
    acc, digest = None, None
    for lo, hi in shard.split_blocks(len(blocks), world, bs, total):
        if hi == lo:
            continue
        pos = lo * bs
        size = min(total, hi * bs) - pos
        d_out = torch.zeros(size + 16, dtype=torch.uint8, device="cuda")
        st = ctx.unpack_blocks_device(d_arch, len(arch), d_out, size, np.ascontiguousarray(blocks[lo:hi]), bs, size)
        assert st == 0
        acc, dg = ctx.blocks_digest(acc, pos, total, d_out)
        out[pos:pos + size] = d_out[:size].cpu().numpy()
        if dg is not None:
            digest = dg
    return out, digest


@pytest.mark.parametrize("total", [1, 100, 240, 241, 5000, 65536, 65537, 3 * 65536, 5 * 65536 + 1234,
                                   4 * 65536 + 1024, 4 * 65536 + 1023, 4 * 65536 + 1025, 2 << 20])
@pytest.mark.parametrize("world", [1, 3])
def test_blocks_path_matches_oracle(gpu_ctx, oracle, total, world):
    data = corpus.big_entry(total, piece=2 * 65536, first=1)        # text | runs | records | random | ...
    frame = oracle.lz4f_encode_port(data, 0, independent=True)
    rc, ref_out, _ = oracle.read_entry_port(2, frame, total, total, oracle.xxh3_port(data))
    assert rc == 0 and np.array_equal(ref_out, data)
    out, digest = _decode_sharded(gpu_ctx, frame, total, world, archive_off=10 + 3 * (total & 7))
    assert np.array_equal(out, data)
    assert digest == oracle.xxh3_port(data)


def test_reference_written_independent_frames(gpu_ctx, oracle, lz4_cases):
    """Frames written by the unmodified reference's LZ4F with blockMode = independent (tests/golden)."""
    seen = 0
    for name, comp in lz4_cases.items():
        if not name.endswith("__indep"):
            continue
        data = lz4_cases[name.replace("__indep", "__in")]
        if zlib.lz4_frame_index(comp) is None or len(data) == 0:
            continue
        for world in (1, 2):
            out, digest = _decode_sharded(gpu_ctx, comp, len(data), world)
            assert np.array_equal(out, data), name
            assert digest == oracle.xxh3_port(data), name
        seen += 1
    assert seen >= 8


def test_declines_a_malformed_block_and_the_entry_path_gives_the_reference_verdict(gpu_ctx, oracle):
    import torch
    total = 3 * 65536
    data = corpus.big_entry(total, piece=65536, first=1)            # text, runs, records
    frame = oracle.lz4f_encode_port(data, 0, independent=True)
    blocks, bs, _ = zlib.lz4_frame_index(frame)
    bad = frame.copy()
    b0 = int(blocks["src_off"][0])
    p = b0 + 1
    lit = int(bad[b0]) >> 4
    if lit == 15:
        while True:
            lit += int(bad[p])
            p += 1
            if bad[p - 1] != 255:
                break
    bad[p + lit] = 0
    bad[p + lit + 1] = 0                                            # first match offset := 0 (lz4.c:2093 rejects)
    d_arch = torch.from_numpy(bad).cuda()
    d_out = torch.zeros(total, dtype=torch.uint8, device="cuda")
    assert gpu_ctx.unpack_blocks_device(d_arch, len(bad), d_out, total, blocks, bs, total) == zlib.ST_NOT_AVAILABLE
    with pytest.raises(zlib.ZpbError):
        gpu_ctx.blocks_digest(None, 0, total, d_out)                # nothing to chain after a declined shard
    # the whole entry through zpb_unpack_device: the class the reference reports (lib/zpack_read.c:421-426)
    e = np.zeros(1, zlib.Entry)
    e["comp_size"], e["dst_cap"], e["uncomp_size"], e["method"] = len(bad), total, total, 2
    status, _ = gpu_ctx.unpack_device(d_arch, len(bad), d_out, total, e)
    rc, _, _ = oracle.read_entry_port(2, bad, total, total, 0)
    assert status[0] == rc == zlib.ST_DECOMPRESS_FAILED


def test_internal_block_method_is_not_reachable_through_the_entry_api(gpu_ctx):
    import torch
    d_arch = torch.zeros(1024, dtype=torch.uint8, device="cuda")
    d_out = torch.zeros(1024, dtype=torch.uint8, device="cuda")
    e = np.zeros(1, zlib.Entry)
    e["comp_size"], e["dst_cap"], e["uncomp_size"], e["method"] = 100, 100, 100, 0x102
    status, _ = gpu_ctx.unpack_device(d_arch, 1024, d_out, 1024, e)
    assert status[0] == zlib.ST_METHOD_INVALID                      # lib/zpack_read.c:459-461


def test_argument_contract(gpu_ctx, oracle):
    import torch
    total = 2 * 65536 + 10
    data = corpus.big_entry(total, piece=65536, first=1)
    frame = oracle.lz4f_encode_port(data, 0, independent=True)
    blocks, bs, _ = zlib.lz4_frame_index(frame)
    d_arch = torch.from_numpy(frame).cuda()
    d_out = torch.zeros(total, dtype=torch.uint8, device="cuda")
    with pytest.raises(zlib.ZpbError):                               # size must match the block count
        gpu_ctx.unpack_blocks_device(d_arch, len(frame), d_out, total, blocks, bs, 65536)
    assert gpu_ctx.unpack_blocks_device(d_arch, len(frame), d_out, total, blocks, bs, total) == 0
    with pytest.raises(zlib.ZpbError):                               # the last shard cannot be 10 bytes on its own
        gpu_ctx.blocks_digest(None, 0, total + 65536, d_out)
    _, dg = gpu_ctx.blocks_digest(None, 0, total, d_out)
    assert dg == oracle.xxh3_port(data)


def test_large_entry_properties(gpu_ctx, oracle):
    """256 MiB entry (4096 blocks): digest agrees with the oracle and with the independent xxh3 kernel over
    the decoded bytes; 4 shards relayed give the same digest as 1."""
    import torch
    total = 256 << 20
    pieces = [corpus.big_entry_piece(k, 1 << 20, total) for k in range(total >> 20)]
    frames = [oracle.lz4f_encode_port(p, 0, independent=True) for p in pieces[:8]]
    # the frame of the whole entry = header + every piece's blocks + EndMark (blocks are independent);
    # only 8 distinct pieces are compressed on the CPU, the entry cycles through them
    body = [f[7:-4] for f in frames]
    frame = np.concatenate([frames[0][:7]] + [body[k & 7] for k in range(total >> 20)] + [np.zeros(4, np.uint8)])
    data = np.concatenate([pieces[k & 7] for k in range(total >> 20)])
    want = oracle.xxh3_port(data)
    for world in (1, 4):
        out, digest = _decode_sharded(gpu_ctx, frame, total, world, archive_off=10)
        assert digest == want
        assert np.array_equal(out, data)
    d = torch.from_numpy(data).cuda()
    assert int(gpu_ctx.xxh3_device(d, [0], [total])[0]) == want


@pytest.mark.parametrize("total", [5 << 20, (9 << 20) + 777])
def test_entry_blocks_host_pipeline(gpu_ctx, oracle, total):
    """zpb_unpack_entry_blocks_host: host buffers, chunks pipelined over the worker streams with the XXH3 state
    relayed chunk to chunk (1 MiB chunks here so that several chunks and several workers are exercised)."""
    import os
    data = corpus.big_entry(total, piece=1 << 20, first=1)
    frame = oracle.lz4f_encode_port(data, 0, independent=True)
    want = oracle.xxh3_port(data)
    os.environ["ZPB_HOST_CHUNK_MB"] = "1"
    import zpack_b200
    ctx = zpack_b200.Context(0)
    del os.environ["ZPB_HOST_CHUNK_MB"]
    try:
        out = np.zeros(total + 5, np.uint8)
        assert ctx.unpack_entry_blocks_host(frame, len(frame), out, total + 5, total, want) == (0, want)
        assert np.array_equal(out[:total], data) and not out[total:].any()
        assert ctx.unpack_entry_blocks_host(frame, len(frame), out, total, total, want ^ 1) == (zlib.ST_HASH_MISMATCH, want)
        assert ctx.unpack_entry_blocks_host(frame, len(frame), out, total - 1, total, want)[0] == zlib.ST_BUFFER_TOO_SMALL
        out[:] = 0
        assert ctx.unpack_entry_blocks_host(frame, len(frame), out, total, total, want, flags=zlib.F_DISCARD) == (0, want)
        assert not out.any()                                        # verdict only: nothing copied back
        # not eligible -> None (the caller goes through zpb_unpack_host): linked frame, wrong declared size
        linked = oracle.lz4f_encode_port(data, 0, independent=False)
        assert ctx.unpack_entry_blocks_host(linked, len(linked), out, total, total, want) is None
        assert ctx.unpack_entry_blocks_host(frame, len(frame), out, total, total - 70000, want) is None
    finally:
        ctx.close()
