"""GPU parity tests proper: the CUDA path, called through the C-ABI, against the oracle on the
same seeded inputs; golden fixtures; reference-side error classes; full-size properties."""
import os

import numpy as np
import pytest

from zpack_b200 import container, corpus
from zpack_b200 import lib as zlib

pytestmark = pytest.mark.gpu

GOLD_HASHES = [0x7874CBA47D02B07D, 0x15F25C0F24DD8E52]


@pytest.fixture(autouse=True, params=["fast", "general"])
def _decode_path(request, gpu_ctx):
    """Every test runs twice: through the scan/parse/exec pipeline (default) and with every entry
    forced through the general decoder."""
    gpu_ctx.set_fast_path(request.param == "fast")
    yield
    gpu_ctx.set_fast_path(True)


def _unpack_all(ctx, arch, d, cap_extra=0, host=True):
    caps = d.uncomp_size + np.uint64(cap_extra)
    e = d.entries(dst_cap=caps)
    out_size = int((e["dst_off"] + e["dst_cap"]).max()) if len(e) else 0
    if host:
        out = np.zeros(max(out_size, 1), np.uint8)
        status, digest = ctx.unpack_host(arch, len(arch), out, out_size, e)
    else:
        import torch
        d_arch = torch.from_numpy(np.ascontiguousarray(arch)).cuda()
        d_out = torch.zeros(max(out_size, 1), dtype=torch.uint8, device="cuda")
        status, digest = ctx.unpack_device(d_arch, len(arch), d_out, out_size, e)
        out = d_out.cpu().numpy()
    return e, status, digest, out


@pytest.mark.parametrize("kind", ["none", "lz4", "zstd"])
@pytest.mark.parametrize("host", [True, False])
def test_golden_archives(gpu_ctx, golden_dir, kind, host):
    """tests/read_archive.c:23-30: 350-byte buffers, memcmp against the plaintext, digest accepted."""
    arch = np.fromfile(os.path.join(golden_dir, f"archive_{kind}.zpk"), np.uint8)
    d = container.parse(arch)
    e, status, digest, out = _unpack_all(gpu_ctx, arch, d, cap_extra=1, host=host)
    assert list(status) == [0, 0]
    assert list(digest) == GOLD_HASHES
    for i, name in enumerate(d.names):
        want = np.fromfile(os.path.join(golden_dir, name), np.uint8)
        o = int(e["dst_off"][i])
        assert np.array_equal(out[o:o + len(want)], want)


@pytest.mark.parametrize("group", [4, 8, 16, 32])
def test_reference_written_lz4_frames_all_group_sizes(gpu_ctx, lz4_cases, oracle, group):
    """Every LZ4 frame fixture written by the unmodified reference (linked, independent, stored,
    checksummed, 256 KB / 4 MB blocks, HC, accelerated) decodes bit-exactly with matching digest."""
    gpu_ctx.set_tuning(group_lanes=group)
    keys = [k for k in lz4_cases if not k.endswith("__in")]
    datas = [lz4_cases[k.split("__")[0] + "__in"] for k in keys]
    hashes = [oracle.xxh3_port(x) for x in datas]
    arch = container.assemble(keys, [lz4_cases[k] for k in keys], [len(x) for x in datas], hashes, [2] * len(keys))
    d = container.parse(arch)
    for extra in (0, 77):
        e, status, digest, out = _unpack_all(gpu_ctx, arch, d, cap_extra=extra, host=False)
        assert (status == 0).all(), dict(zip(keys, status))
        assert np.array_equal(digest, np.array(hashes, np.uint64))
        for i, x in enumerate(datas):
            o = int(e["dst_off"][i])
            assert np.array_equal(out[o:o + len(x)], x), keys[i]
    gpu_ctx.set_tuning(group_lanes=8)


def test_xxh3_kat_on_gpu(gpu_ctx, golden_dir):
    import json
    kat = json.load(open(os.path.join(golden_dir, "xxh3_kat.json")))
    g, buf = 2654435761, bytearray(70000)
    for i in range(70000):
        buf[i] = (g >> 56) & 0xFF
        g = (g * 11400714785074694797) & 0xFFFFFFFFFFFFFFFF
    buf = np.frombuffer(bytes(buf), np.uint8)
    import torch
    d = torch.from_numpy(buf.copy()).cuda()
    lens = sorted({int(k) for t in ("upstream", "reference_run") for k in kat[t]})
    want = {int(k): int(v, 16) for t in ("upstream", "reference_run") for k, v in kat[t].items()}
    got = gpu_ctx.xxh3_device(d, [0] * len(lens), lens)
    for n, h in zip(lens, got):
        assert int(h) == want[n], n
    # unaligned starts take the byte-assembled path: same digests as the oracle
    from oracle import oracle as O
    offs = [1, 3, 7, 9, 15, 17]
    got = gpu_ctx.xxh3_device(d, offs, [5000] * len(offs))
    for o, h in zip(offs, got):
        assert int(h) == O.xxh3_port(buf[o:o + 5000])
    assert gpu_ctx.xxh3_host(buf[:12345]) == O.xxh3_port(buf[:12345])
    assert gpu_ctx.xxh3_host(b"") == 0x2D06800538D394C2


def test_seeded_corpus_vs_oracle_mixed_sizes(gpu_ctx, oracle):
    """Differential decode: ragged sizes incl. 0, 1, block boundaries; linked + independent frames."""
    sizes = [0, 1, 5, 12, 13, 64, 240, 241, 1023, 1024, 1025, 4096, 65535, 65536, 65537, 131072, 200001, 262144]
    bufs, names, payload = [], [], []
    for i, s in enumerate(sizes * 2):
        b = corpus.entry_bytes(i, s)
        bufs.append(b)
        names.append(f"f{i}")
        payload.append(oracle.lz4f_encode_port(b, 0, independent=bool(i & 1)))
    hashes = [oracle.xxh3_port(b) for b in bufs]
    arch = container.assemble(names, payload, [len(b) for b in bufs], hashes, [2] * len(bufs))
    d = container.parse(arch)
    for host in (True, False):
        e, status, digest, out = _unpack_all(gpu_ctx, arch, d, host=host)
        assert (status == 0).all(), status
        assert np.array_equal(digest, np.array(hashes, np.uint64))
        for i, b in enumerate(bufs):
            rc, want, dg = oracle.read_entry_port(2, payload[i], len(b), len(b), hashes[i])
            assert rc == 0
            o = int(e["dst_off"][i])
            assert np.array_equal(out[o:o + len(b)], want), i


def test_error_classes_match_oracle(gpu_ctx, oracle):
    """Truncated / bit-flipped / lying entries: same zpack_result class as the reference path,
    and one bad entry never poisons its neighbours."""
    rng = np.random.default_rng(5)
    size = 50000
    bufs = [corpus.entry_bytes(i, size) for i in range(8)]
    payload = [oracle.lz4f_encode_port(b, 0, independent=False) for b in bufs]
    hashes = [oracle.xxh3_port(b) for b in bufs]
    arch = container.assemble([f"f{i}" for i in range(8)], payload, [size] * 8, hashes, [2] * 8)
    d = container.parse(arch)
    e = d.entries()
    cases = []
    for trial in range(60):
        m = arch.copy()
        i = int(rng.integers(0, 8))
        lo, hi = int(d.offset[i]), int(d.offset[i] + d.comp_size[i])
        pos = int(rng.integers(lo, hi))
        m[pos] ^= 1 << int(rng.integers(0, 8))
        cases.append((m, i, lo, hi))
    for m, i, lo, hi in cases:
        out = np.zeros(int((e["dst_off"] + e["dst_cap"]).max()), np.uint8)
        status, digest = gpu_ctx.unpack_host(m, len(m), out, len(out), e)
        rc, _, _ = oracle.read_entry_port(2, m[lo:hi], size, size, hashes[i])
        assert (status[i] == 0) == (rc == 0), (i, status[i], rc)
        if rc != 0:
            assert status[i] in (12, 13, 15, 17), status[i]
        others = [k for k in range(8) if k != i]
        assert (status[others] == 0).all()
    # lying descriptors
    e2 = e.copy()
    e2["hash"][0] ^= np.uint64(1)            # wrong digest          -> HASH_MISMATCH, buffer still filled
    e2["dst_cap"][1] = size - 1              # max_size < uncomp     -> BUFFER_TOO_SMALL (zpack_read.c:329)
    e2["comp_size"][2] -= np.uint64(9)       # truncated entry       -> FILE_INCOMPLETE
    e2["method"][3] = 7                      # unknown method        -> COMP_METHOD_INVALID
    e2["comp_size"][4] = 0                   # empty entry           -> OK, untouched, no hash (zpack_read.c:328)
    e2["src_off"][5] = len(arch) + 5         # outside the archive   -> FILE_OFFSET_INVALID
    out = np.zeros(int((e["dst_off"] + e["dst_cap"]).max()), np.uint8)
    status, digest = gpu_ctx.unpack_host(arch, len(arch), out, len(out), e2)
    assert list(status[:6]) == [15, 12, 17, 19, 0, 16], status
    assert np.array_equal(out[:size], bufs[0])
    assert (status[6:] == 0).all()


def test_multi_frame_and_skippable_frames(gpu_ctx, oracle):
    a, b = corpus.entry_bytes(1, 30000), corpus.entry_bytes(2, 90000)
    skip = np.frombuffer(bytes([0x53, 0x2A, 0x4D, 0x18, 5, 0, 0, 0]) + b"hello", np.uint8)
    comp = np.concatenate([oracle.lz4f_encode_port(a), skip, oracle.lz4f_encode_port(b, independent=True)])
    whole = np.concatenate([a, b])
    h = oracle.xxh3_port(whole)
    rc, want, _ = oracle.read_entry_port(2, comp, len(whole), len(whole), h)
    assert rc == 0 and np.array_equal(want, whole)
    arch = container.assemble(["x"], [comp], [len(whole)], [h], [2])
    d = container.parse(arch)
    e, status, digest, out = _unpack_all(gpu_ctx, arch, d)
    assert list(status) == [0] and int(digest[0]) == h and np.array_equal(out[:len(whole)], whole)


def test_reference_written_archive_if_ref_present(gpu_ctx, oracle):
    """The drop-in case: an archive packed by the UNMODIFIED reference (zpack_write_archive,
    linked 64 KB blocks) unpacks bit-exactly on the GPU."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not present")
    n, size = 256, 131072
    bufs = [corpus.entry_bytes(i, size) for i in range(n)]
    arch = oracle.write_archive_ref([corpus.entry_name(i) for i in range(n)], bufs, 2, 0)
    d = container.parse(arch)
    e, status, digest, out = _unpack_all(gpu_ctx, arch, d, host=False)
    assert (status == 0).all()
    assert np.array_equal(digest, d.hash)
    assert np.array_equal(out.reshape(n, size), np.stack(bufs))


def test_full_size_property_roundtrip(gpu_ctx, oracle):
    """BASELINE C2-shaped slice at full entry size: digests of every entry equal the packer's,
    and a checksum of checksums over the whole output equals the CPU one."""
    n, size = 2048, 131072
    import torch
    bufs = [corpus.entry_bytes(i, size) for i in range(n)]
    payload = [oracle.lz4f_encode_port(b, 0, independent=bool((i >> 2) & 1)) for i, b in enumerate(bufs)]
    hashes = np.array([oracle.xxh3_port(b) for b in bufs], np.uint64)
    arch = container.assemble([f"{i}" for i in range(n)], payload, [size] * n, hashes, [2] * n)
    d = container.parse(arch)
    e, status, digest, out = _unpack_all(gpu_ctx, arch, d, host=False)
    assert (status == 0).all()
    assert np.array_equal(digest, hashes)
    assert oracle.xxh3_port(out[:n * size]) == oracle.xxh3_port(np.concatenate(bufs))


def test_pipelined_host_path_and_discard_flag(oracle, monkeypatch):
    """zpb_unpack_host cuts big batches into chunks dealt to worker sub-contexts (H2D / kernels / D2H of
    different chunks overlap).  Force tiny chunks so that a test-sized batch takes that path; results must be
    identical to the single-shot path, in the caller's entry order.  ZPB_F_DISCARD keeps the bytes on the
    device (the `zpack t` case): verdicts and digests still come back, the host buffer stays untouched."""
    import zpack_b200
    monkeypatch.setenv("ZPB_HOST_CHUNK_MB", "1")
    monkeypatch.setenv("ZPB_HOST_WORKERS", "3")
    ctx = zpack_b200.Context(0)
    try:
        n = 96
        sizes = [131072 if i % 5 else 70001 for i in range(n)]
        bufs = [corpus.entry_bytes(i, s) for i, s in enumerate(sizes)]
        payload = [oracle.lz4f_encode_port(b, 0, independent=bool(i & 1)) for i, b in enumerate(bufs)]
        hashes = np.array([oracle.xxh3_port(b) for b in bufs], np.uint64)
        arch = container.assemble([f"{i}" for i in range(n)], payload, sizes, hashes, [2] * n)
        d = container.parse(arch)
        e = d.entries()
        e["hash"][7] ^= np.uint64(1)
        perm = np.random.default_rng(3).permutation(n)          # caller order != archive order
        ep = np.ascontiguousarray(e[perm])
        out_size = int((e["dst_off"] + e["dst_cap"]).max())
        out = np.zeros(out_size, np.uint8)
        status, digest = ctx.unpack_host(arch, len(arch), out, out_size, ep)
        want = np.zeros(n, np.int32)
        want[7] = 15
        assert np.array_equal(status, want[perm])
        assert np.array_equal(digest, hashes[perm])
        for i, b in enumerate(bufs):
            o = int(e["dst_off"][i])
            assert np.array_equal(out[o:o + len(b)], b), i
        ev = ep.copy()
        ev["flags"] |= zlib.F_DISCARD
        out2 = np.full(out_size, 0xEE, np.uint8)
        status, digest = ctx.unpack_host(arch, len(arch), out2, out_size, ev)
        assert np.array_equal(status, want[perm]) and np.array_equal(digest, hashes[perm])
        assert (out2 == 0xEE).all()
    finally:
        ctx.close()


def test_misaligned_output_slot_is_refused_not_a_fault(gpu_ctx, oracle):
    """include/zpack_b200.h documents dst_off as 16-byte aligned (the kernels store 16 bytes at a time).  An odd slot on the
    device path must come back as ZPB_E_ARG — and leave the context usable — instead of a sticky misaligned-address fault."""
    import torch
    import zpack_b200
    bufs = [corpus.entry_bytes(i, 40000) for i in range(4)]
    frames = [oracle.lz4f_encode_port(b, 0, independent=False) for b in bufs]
    hashes = [oracle.xxh3_port(b) for b in bufs]
    arch = container.assemble([f"f{i}" for i in range(4)], frames, [40000] * 4, hashes, [2] * 4)
    e = container.parse(arch).entries()
    d_arch = torch.from_numpy(arch).cuda()
    d_out = torch.zeros(int((e["dst_off"] + e["dst_cap"]).max()) + 64, dtype=torch.uint8, device="cuda")
    bad = e.copy()
    bad["dst_off"][2] += 3
    with pytest.raises(zpack_b200.lib.ZpbError, match="16-byte aligned"):
        gpu_ctx.unpack_device(d_arch, len(arch), d_out, d_out.numel(), bad)
    st, dg = gpu_ctx.unpack_device(d_arch, len(arch), d_out, d_out.numel(), e)     # the context is still good
    assert (st == 0).all() and np.array_equal(dg, np.array(hashes, np.uint64))
