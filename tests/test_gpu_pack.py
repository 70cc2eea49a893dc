"""GPU pack path (zpb_pack_*): GPU-written LZ4 frames must be valid for the oracle's frame decoder and for
the UNMODIFIED reference reader, round-trip bit-exactly, carry the right XXH3-64, and compress about as
well as the reference at the same level (compressed bytes themselves are unpinned: tests/write_archive.c
only checks return codes)."""
import os
import subprocess

import numpy as np
import pytest

from zpack_b200 import container, corpus
from zpack_b200 import lib as zlib

pytestmark = pytest.mark.gpu

SIZES = [0, 1, 5, 12, 13, 14, 64, 1000, 4095, 65535, 65536, 65537, 131072, 200001, 262144]


def _files(ctx, sizes, method=2, level=0, first=0):
    bufs = [corpus.entry_bytes(first + i, s) for i, s in enumerate(sizes)]
    f = np.zeros(len(bufs), zlib.File)
    in_off = out_off = 0
    for i, b in enumerate(bufs):
        f["src_off"][i], f["size"][i] = in_off, len(b)
        cap = ctx.pack_bound(method, len(b))
        f["dst_off"][i], f["dst_cap"][i] = out_off, cap
        f["method"][i], f["level"][i] = method, level
        in_off += (len(b) + 15) & ~15
        out_off += (cap + 15) & ~15
    h_in = np.zeros(max(in_off, 16), np.uint8)
    for i, b in enumerate(bufs):
        h_in[int(f["src_off"][i]):int(f["src_off"][i]) + len(b)] = b
    return bufs, f, h_in, max(out_off, 16)


def _pack(ctx, sizes, host=True, **kw):
    bufs, f, h_in, out_size = _files(ctx, sizes, **kw)
    if host:
        h_out = np.zeros(out_size, np.uint8)
        comp, digest, status = ctx.pack_host(h_in, len(h_in), h_out, out_size, f)
    else:
        import torch
        d_in = torch.from_numpy(h_in).cuda()
        d_out = torch.zeros(out_size, dtype=torch.uint8, device="cuda")
        comp, digest, status = ctx.pack_device(d_in, len(h_in), d_out, out_size, f)
        h_out = d_out.cpu().numpy()
    frames = [h_out[int(f["dst_off"][i]):int(f["dst_off"][i] + comp[i])].copy() for i in range(len(bufs))]
    return bufs, frames, comp, digest, status


@pytest.mark.parametrize("host", [True, False])
def test_gpu_frames_decode_with_the_oracle(gpu_ctx, oracle, host):
    bufs, frames, comp, digest, status = _pack(gpu_ctx, SIZES * 2, host=host)
    assert (status == 0).all(), status
    for i, (b, fr) in enumerate(zip(bufs, frames)):
        assert len(fr) <= gpu_ctx.pack_bound(2, len(b))
        assert bytes(fr[:6]) == bytes([0x04, 0x22, 0x4D, 0x18, 0x60, 0x40])   # B.Indep frames, 64 KB blocks
        rc, out = oracle.lz4f_decode_port(fr, len(b))
        assert rc == 0 and np.array_equal(out, b), (i, len(b))
        assert int(digest[i]) == oracle.xxh3_port(b)
    assert len(frames[0]) == 11                                                # empty file: header + EndMark


def test_gpu_written_archive_round_trips_through_the_reference_reader(gpu_ctx, oracle, tmp_path):
    """The validity gate of BASELINE config C3: reference zpack_read_file (and its CLI `t`) accept every entry."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not present")
    n, size = 128, 131072
    bufs, frames, comp, digest, status = _pack(gpu_ctx, [size] * n, host=False)
    assert (status == 0).all()
    names = [corpus.entry_name(i) for i in range(n)]
    arch = container.assemble(names, frames, [size] * n, digest, [2] * n)       # offsets: host prefix sum
    rd = oracle.RefReader(arch)
    assert rd.count == n
    for i in range(n):
        rc, out = rd.read(i)
        assert rc == 0, (i, rc)                                                 # decode OK and XXH3 verified by the reference
        assert np.array_equal(out, bufs[i])
    rd.close()
    p = tmp_path / "gpu.zpk"
    arch.tofile(p)
    r = subprocess.run([oracle.REF_CLI, "t", str(p)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"Corrupted files: 0/{n}" in r.stdout, r.stdout[-300:]


def test_ratio_against_the_reference_level0(gpu_ctx, oracle):
    n, size = 256, 131072
    bufs, frames, comp, digest, status = _pack(gpu_ctx, [size] * n, host=False)
    assert (status == 0).all()
    ours = float(comp.sum())
    ref = float(sum(len(oracle.lz4f_encode_port(b, 0, independent=True)) for b in bufs))
    ratio_ours, ratio_ref = n * size / ours, n * size / ref
    print(f"\nLZ4 level-0 ratio on zpk-synth-v1: GPU {ratio_ours:.3f} vs reference-algorithm {ratio_ref:.3f}")
    assert ratio_ours > 0.85 * ratio_ref


def test_zstd_ratio_against_the_reference_level3(gpu_ctx, oracle):
    """Every class of the corpus through the zstd writer: frames decode in the oracle's port (and in ZSTD_decompress),
    ratio printed next to ZSTD_compress level 3 (the match finder is the LZ4 block compressor's, literals are raw)."""
    n, size = 64, 131072
    bufs, frames, comp, digest, status = _pack(gpu_ctx, [size] * n, host=False, method=1)
    assert (status == 0).all()
    for i in range(0, n, 5):
        rc, got = oracle.zstd_decode_port(frames[i], size)
        assert rc == 0 and np.array_equal(got[:size], bufs[i]), i
        if oracle.have_ref():
            assert np.array_equal(oracle.zstd_decompress_ref(frames[i], size), bufs[i])
    ours = n * size / float(comp.sum())
    lz4 = n * size / float(_pack(gpu_ctx, [size] * n, host=False)[2].sum())
    msg = f"\nzstd writer ratio on zpk-synth-v1: GPU {ours:.3f} (its LZ4 frames: {lz4:.3f})"
    if oracle.have_ref():
        ref = n * size / float(sum(len(oracle.zstd_compress_ref(b, 3)) for b in bufs))
        msg += f" vs ZSTD_compress level 3 {ref:.3f}"
    print(msg)
    assert ours > lz4


def test_pack_then_unpack_on_gpu(gpu_ctx):
    sizes = [131072] * 64 + SIZES
    bufs, frames, comp, digest, status = _pack(gpu_ctx, sizes, host=False, first=100)
    assert (status == 0).all()
    arch = container.assemble([f"f{i}" for i in range(len(bufs))], frames, [len(b) for b in bufs], digest, [2] * len(bufs))
    d = container.parse(arch)
    e = d.entries()
    out = np.zeros(int((e["dst_off"] + e["dst_cap"]).max()), np.uint8)
    st, dg = gpu_ctx.unpack_host(arch, len(arch), out, len(out), e)
    assert (st == 0).all(), st
    for i, b in enumerate(bufs):
        o = int(e["dst_off"][i])
        assert np.array_equal(out[o:o + len(b)], b), i


def test_none_method_and_error_statuses(gpu_ctx, oracle):
    bufs, frames, comp, digest, status = _pack(gpu_ctx, [0, 7, 70000], method=0)
    assert (status == 0).all()
    for b, fr, dg in zip(bufs, frames, digest):
        assert np.array_equal(fr, b) and int(dg) == oracle.xxh3_port(b)
    # zstd: frames of 64 KB Compressed_Blocks built from the block compressor's matches (raw literals + predefined-mode
    # FSE sequences, zstd_encode.cuh), Raw_Blocks where that does not pay; the oracle's port, the unmodified reference
    # decoder and our GPU decoder must all read them back
    sizes = [0, 1, 1000, 131072, 131073, 400000]
    bufs, frames, comp, digest, status = _pack(gpu_ctx, sizes, method=1)
    assert (status == 0).all(), status
    assert all(len(fr) <= gpu_ctx.pack_bound(1, len(b)) for fr, b in zip(frames, bufs))
    assert comp[3] < 0.8 * sizes[3] and comp[5] < 0.8 * sizes[5]                  # it does compress (records / text classes)
    for b, fr, dg in zip(bufs, frames, digest):
        assert int(dg) == oracle.xxh3_port(b)
        rc, got = oracle.zstd_decode_port(fr, len(b))
        assert rc == 0 and np.array_equal(got[:len(b)], b)
        if oracle.have_ref():
            assert np.array_equal(oracle.zstd_decompress_ref(fr, len(b)), b)
    arch = container.assemble([f"z{i}" for i in range(len(bufs))], frames, [len(b) for b in bufs], digest, [1] * len(bufs))
    e = container.parse(arch).entries()
    out = np.zeros(int((e["dst_off"] + e["dst_cap"]).max()) + 16, np.uint8)
    st, dg2 = gpu_ctx.unpack_host(arch, len(arch), out, len(out), e)
    assert (st == 0).all(), st
    _, _, _, _, status = _pack(gpu_ctx, [1000], method=2, level=9)   # LZ4 HC levels: not built, never a silent downgrade
    assert list(status) == [24]
    _, _, _, _, status = _pack(gpu_ctx, [1000], method=9)
    assert list(status) == [19]
    bufs, f, h_in, out_size = _files(gpu_ctx, [5000])
    f["dst_cap"][0] = 100                                            # slot smaller than the bound
    comp, digest, status = gpu_ctx.pack_host(h_in, len(h_in), np.zeros(out_size, np.uint8), out_size, f)
    assert list(status) == [14]


def test_pipelined_pack_host_matches_the_device_path(oracle, monkeypatch):
    """zpb_pack_host cuts big batches into chunks over worker sub-contexts and brings each chunk's frames back with
    one gathered D2H.  Tiny chunks force that path on a test-sized batch; caller order != source order, slots at
    odd (not 16-byte aligned) offsets, zero-length and mixed-method files included.  Frames, sizes and digests must
    equal what the single-launch device path produces, and slack between slots must stay untouched."""
    import torch
    import zpack_b200
    monkeypatch.setenv("ZPB_HOST_CHUNK_MB", "1")
    monkeypatch.setenv("ZPB_HOST_WORKERS", "3")
    ctx = zpack_b200.Context(0)
    try:
        n = 80
        sizes = [0 if i == 11 else (131072 if i % 3 else 50001 + i) for i in range(n)]
        bufs = [corpus.entry_bytes(i, s) for i, s in enumerate(sizes)]
        f = np.zeros(n, zlib.File)
        in_off, out_off = 0, 3
        for i, b in enumerate(bufs):
            m = 0 if i % 7 == 5 else (1 if i % 7 == 3 else 2)     # stored, zstd and LZ4 files in one batch
            f["src_off"][i], f["size"][i] = in_off, len(b)
            cap = ctx.pack_bound(m, len(b))
            f["dst_off"][i], f["dst_cap"][i] = out_off, cap
            f["method"][i] = m
            in_off += (len(b) + 15) & ~15
            out_off += cap + 5                       # odd slot starts
        h_in = np.zeros(in_off + 16, np.uint8)
        for i, b in enumerate(bufs):
            h_in[int(f["src_off"][i]):int(f["src_off"][i]) + len(b)] = b
        perm = np.random.default_rng(5).permutation(n)
        fp = np.ascontiguousarray(f[perm])
        out_size = out_off + 16
        h_out = np.full(out_size, 0xA5, np.uint8)
        comp, digest, status = ctx.pack_host(h_in, len(h_in), h_out, out_size, fp)
        d_in = torch.from_numpy(h_in).cuda()
        d_out = torch.zeros(out_size, dtype=torch.uint8, device="cuda")
        comp_d, digest_d, status_d = ctx.pack_device(d_in, len(h_in), d_out, out_size, fp)
        ref = d_out.cpu().numpy()
        assert (status == 0).all() and np.array_equal(status, status_d)
        assert np.array_equal(comp, comp_d) and np.array_equal(digest, digest_d)
        touched = np.zeros(out_size, bool)
        for k in range(n):
            o, c = int(fp["dst_off"][k]), int(comp[k])
            assert np.array_equal(h_out[o:o + c], ref[o:o + c]), k
            assert digest[k] == oracle.xxh3_port(bufs[perm[k]])
            if fp["method"][k] == 1 and k % 3 == 0:     # the zstd frames of a mixed batch decode in the CPU checker
                rc, got = oracle.zstd_decode_port(h_out[o:o + c], len(bufs[perm[k]]))
                assert rc == 0 and np.array_equal(got[:len(bufs[perm[k]])], bufs[perm[k]]), k
            touched[o:o + c] = True
        assert (h_out[~touched] == 0xA5).all()       # nothing outside the frames was written on the host side
    finally:
        ctx.close()


@pytest.mark.parametrize("method", [2, 1])
def test_files_of_varied_statistics_round_trip(gpu_ctx, oracle, method):
    """Alphabets from 2 to 256 symbols, skewed symbols, repeats at distances from 1 to 60 000, sizes that are not block or
    window multiples: LZ4 and zstd frames from the GPU writer must decode in the CPU checker and on the GPU reader."""
    rng = np.random.default_rng(123 + method)
    bufs = []
    for t in range(40):
        n = int(rng.choice([1, 12, 13, 4095, 4109, 65535, 65536, 65537, 100000, 200000, 262144]))
        alpha = int(rng.choice([2, 5, 16, 64, 120, 256]))
        b = rng.integers(0, alpha, n, dtype=np.uint8) if t % 3 else (rng.zipf(1.3, n) % alpha).astype(np.uint8)
        for _ in range(int(rng.choice([0, 50, 400, 3000]))):
            ln = int(rng.choice([4, 5, 8, 20, 70, 300]))
            if n <= 2 * ln + 2:
                break
            p = int(rng.integers(ln + 1, n - ln))
            d = min(int(rng.choice([1, 2, 3, 7, 64, 1000, 60000])), p)
            for k in range(ln):
                b[p + k] = b[p + k - d]
        bufs.append(b)
    f = np.zeros(len(bufs), zlib.File)
    in_off = out_off = 0
    for i, b in enumerate(bufs):
        cap = gpu_ctx.pack_bound(method, len(b))
        f["src_off"][i], f["size"][i], f["dst_off"][i], f["dst_cap"][i], f["method"][i] = in_off, len(b), out_off, cap, method
        in_off += (len(b) + 15) & ~15
        out_off += (cap + 15) & ~15
    h_in = np.zeros(in_off + 16, np.uint8)
    for i, b in enumerate(bufs):
        h_in[int(f["src_off"][i]):int(f["src_off"][i]) + len(b)] = b
    h_out = np.zeros(out_off + 16, np.uint8)
    comp, digest, status = gpu_ctx.pack_host(h_in, len(h_in), h_out, len(h_out), f)
    assert (status == 0).all(), status
    frames = [h_out[int(f["dst_off"][i]):int(f["dst_off"][i] + comp[i])].copy() for i in range(len(bufs))]
    for i, (b, fr) in enumerate(zip(bufs, frames)):
        assert int(digest[i]) == oracle.xxh3_port(b)
        rc, got = (oracle.lz4f_decode_port if method == 2 else oracle.zstd_decode_port)(fr, len(b))
        assert rc == 0 and np.array_equal(got[:len(b)], b), (i, len(b))
    arch = container.assemble([f"v{i}" for i in range(len(bufs))], frames, [len(b) for b in bufs], digest, [method] * len(bufs))
    e = container.parse(arch).entries()
    out = np.zeros(int((e["dst_off"] + e["dst_cap"]).max()) + 16, np.uint8)
    st, dg = gpu_ctx.unpack_host(arch, len(arch), out, len(out), e)
    assert (st == 0).all(), st
    for i, b in enumerate(bufs):
        o = int(e["dst_off"][i])
        assert np.array_equal(out[o:o + len(b)], b), i
