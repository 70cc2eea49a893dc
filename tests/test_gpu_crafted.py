"""Hand-built LZ4 frames that aim at specific branches of the K2 execute kernel (lz4_fast.cuh): pattern fills
(offsets 1 / 2 / 4), short periods that are not fills (3, 5, 7 ...), 16-byte-chunk copies from the ring and from
flushed output, matches that cross the 4 KiB ring, redirected chains (a match whose source is an earlier match of
the same 32-sequence step), literal runs on both sides of the lane-copy thresholds, block-linked references into
the previous block.  The expected bytes come from the builder itself (a direct restatement of the sequence
semantics, lz4_Block_format.md) and are cross-checked with the oracle's decoder before the GPU is asked."""
import numpy as np
import pytest

from zpack_b200 import container

pytestmark = pytest.mark.gpu

from crafted import LINKED_HDR, OFFSETS, frame as _frame  # noqa: E402,F401


@pytest.mark.parametrize("fast", [True, False])
def test_crafted_sequences(gpu_ctx, oracle, fast):
    rng = np.random.default_rng(0xC0FFEE)
    frames, plain = [], []
    for style in ("mixed", "chains", "fills", "long"):
        for sizes in ([65536], [65536, 65536], [65536, 65536, 30000], [20000], [65536, 1000]):
            for _ in range(3):
                fr, data = _frame(rng, sizes, style)
                rc, got = oracle.lz4f_decode_port(np.frombuffer(fr, np.uint8), len(data))
                assert rc == 0 and bytes(got) == data, "builder and oracle disagree"
                frames.append(np.frombuffer(fr, np.uint8))
                plain.append(np.frombuffer(data, np.uint8))
    hashes = [oracle.xxh3_port(b) for b in plain]
    arch = container.assemble([f"c{i}" for i in range(len(frames))], frames, [len(b) for b in plain], hashes,
                              [2] * len(frames))
    d = container.parse(arch)
    e = d.entries()
    out_size = int((e["dst_off"] + e["dst_cap"]).max())
    import torch
    gpu_ctx.set_fast_path(fast)
    try:
        d_arch = torch.from_numpy(arch).cuda()
        d_out = torch.zeros(out_size, dtype=torch.uint8, device="cuda")
        status, digest = gpu_ctx.unpack_device(d_arch, len(arch), d_out, out_size, e)
    finally:
        gpu_ctx.set_fast_path(True)
    out = d_out.cpu().numpy()
    assert (status == 0).all(), status
    assert np.array_equal(digest, np.array(hashes, np.uint64))
    for i, b in enumerate(plain):
        o = int(e["dst_off"][i])
        assert np.array_equal(out[o:o + len(b)], b), i
