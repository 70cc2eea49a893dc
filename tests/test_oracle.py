"""Pins the CPU oracle (oracle/*_oracle.c) before anything trusts it: golden vectors of the
reference's own tests, upstream known-answer tables, and the unmodified reference (oracle/_ref)."""
import json
import os

import numpy as np
import pytest

from zpack_b200 import container, corpus

GOLD_HASHES = {"file1.txt": 0x7874CBA47D02B07D, "file2.txt": 0x15F25C0F24DD8E52}  # tests/archive.h:112-115
GOLD_SIZES = {"file1.txt": 169, "file2.txt": 349}                                  # tests/archive.h:107-110


def xxh_buffer(n):
    g, out = 2654435761, bytearray(n)
    for i in range(n):
        out[i] = (g >> 56) & 0xFF
        g = (g * 11400714785074694797) & 0xFFFFFFFFFFFFFFFF
    return bytes(out)


def test_xxh3_upstream_kat_and_reference_runs(oracle, golden_dir):
    kat = json.load(open(os.path.join(golden_dir, "xxh3_kat.json")))
    buf = xxh_buffer(70000)
    for table in ("upstream", "reference_run"):
        for n, h in kat[table].items():
            assert oracle.xxh3_port(buf[:int(n)]) == int(h, 16), (table, n)


def test_xxh3_golden_file_digests(oracle, golden_dir):
    for name, h in GOLD_HASHES.items():
        data = open(os.path.join(golden_dir, name), "rb").read()
        assert len(data) == GOLD_SIZES[name]
        assert oracle.xxh3_port(data) == h


@pytest.mark.parametrize("kind", ["none", "zstd", "lz4"])
def test_golden_archives_open_and_decode(oracle, golden_dir, kind):
    """tests/open_archive.c:21-24 + tests/read_archive.c:23-30 restated against the port."""
    raw = np.fromfile(os.path.join(golden_dir, f"archive_{kind}.zpk"), np.uint8)
    d = container.parse(raw)
    assert d.names == ["file1.txt", "file2.txt"]
    for i, name in enumerate(d.names):
        assert int(d.uncomp_size[i]) == GOLD_SIZES[name]
        assert int(d.hash[i]) == GOLD_HASHES[name]
        comp = raw[int(d.offset[i]):int(d.offset[i] + d.comp_size[i])]
        rc, out, dg = oracle.read_entry_port(int(d.method[i]), comp, 350, int(d.uncomp_size[i]), int(d.hash[i]))
        if kind == "zstd" and rc == 13 and not _zstd_ready(oracle):
            pytest.skip("zstd restatement not built yet")
        assert rc == 0
        assert out.tobytes() == open(os.path.join(golden_dir, name), "rb").read()
        assert dg == GOLD_HASHES[name]


def _zstd_ready(oracle):
    rc, out = oracle.zstd_decode_port(np.array([0x28, 0xB5, 0x2F, 0xFD, 0x20, 0x00, 0x01, 0x00, 0x00], np.uint8), 8)
    return rc == 0


def test_lz4_frames_written_by_reference(oracle, lz4_cases):
    for key, frame in lz4_cases.items():
        if key.endswith("__in"):
            continue
        data = lz4_cases[key.split("__")[0] + "__in"]
        rc, out = oracle.lz4f_decode_port(frame, len(data))
        assert rc == 0, key
        assert np.array_equal(out, data), key
        # with spare room as in tests/read_archive.c (max_size > uncomp_size)
        rc, out = oracle.lz4f_decode_port(frame, len(data) + 77)
        assert rc == 0 and np.array_equal(out, data), key


def test_lz4_error_classes(oracle, lz4_cases):
    data, frame = lz4_cases["text_200k__in"], lz4_cases["text_200k__linked"]
    rc, _ = oracle.lz4f_decode_port(frame[:len(frame) // 2], len(data))
    assert rc == 17                                   # FILE_INCOMPLETE
    rc, _ = oracle.lz4f_decode_port(frame, len(data) - 1)
    assert rc == 12                                   # BUFFER_TOO_SMALL
    bad = frame.copy(); bad[0] ^= 0xFF
    assert oracle.lz4f_decode_port(bad, len(data))[0] == 13
    bad = frame.copy(); bad[6] ^= 0x01                # header check byte
    assert oracle.lz4f_decode_port(bad, len(data))[0] == 13


def test_lz4_port_encoder_is_byte_identical_to_reference_frames(oracle, lz4_cases):
    """The greedy hash-table encoder restated from lz4.c:851-1240 reproduces the reference's
    frames byte for byte (fresh context), in the mode the reference frame declares in its FLG
    byte (LZ4F_compressFrame silently switches inputs <= 64 KB to independent, lz4frame.c:417)."""
    for key, ref_frame in lz4_cases.items():
        if key.endswith("__in") or key.split("__")[1] not in ("linked", "indep"):
            continue
        data = lz4_cases[key.split("__")[0] + "__in"]
        indep = bool(ref_frame[4] & 0x20)
        frame = oracle.lz4f_encode_port(data, 0, indep)
        assert np.array_equal(frame, ref_frame), key
        rc, out = oracle.lz4f_decode_port(frame, len(data))
        assert rc == 0 and np.array_equal(out, data), key


# ---------------------------------------------------------------- differential vs the unmodified reference
def _need_ref(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not present (built only where /root/reference exists)")


def test_ref_agrees_on_xxh3_random_lengths(oracle):
    _need_ref(oracle)
    rng = np.random.default_rng(7)
    buf = rng.integers(0, 256, 300000, dtype=np.uint8)
    for n in list(range(0, 300)) + [int(x) for x in rng.integers(300, 300000, 60)]:
        assert oracle.xxh3_port(buf[:n]) == oracle.xxh3_ref(buf[:n]), n


def test_ref_reader_accepts_port_written_frames_and_vice_versa(oracle):
    """Archives built from port-encoded frames open + verify in the reference reader."""
    _need_ref(oracle)
    n, size = 16, 70001
    bufs = [corpus.entry_bytes(i, size) for i in range(n)]
    names = [corpus.entry_name(i) for i in range(n)]
    payload = [oracle.lz4f_encode_port(b, 0, independent=(i & 4) != 0) for i, b in enumerate(bufs)]
    hashes = [oracle.xxh3_port(b) for b in bufs]
    arch = container.assemble(names, payload, [size] * n, hashes, [2] * n)
    r = oracle.RefReader(arch)
    assert r.count == n
    for i in range(n):
        rc, out = r.read(i)
        assert rc == 0 and np.array_equal(out, bufs[i]), i
    r.close()
    # and the other direction: reference-written archive, decoded by the port
    arch2 = oracle.write_archive_ref(names, bufs, 2, 0)
    d = container.parse(arch2)
    assert d.names == names
    for i in range(n):
        comp = arch2[int(d.offset[i]):int(d.offset[i] + d.comp_size[i])]
        rc, out, dg = oracle.read_entry_port(2, comp, size, size, int(d.hash[i]))
        assert rc == 0 and np.array_equal(out, bufs[i]) and dg == hashes[i]


def test_ref_and_port_agree_on_corrupted_lz4_entries(oracle):
    """Adversarial parity: same result class for truncated / bit-flipped entries."""
    _need_ref(oracle)
    rng = np.random.default_rng(11)
    size = 40000
    bufs = [corpus.entry_bytes(i, size) for i in range(1, 4)]
    names = ["a", "b", "c"]
    arch = oracle.write_archive_ref(names, bufs, 2, 0)
    d = container.parse(arch)
    for i in range(3):
        lo, hi = int(d.offset[i]), int(d.offset[i] + d.comp_size[i])
        for trial in range(40):
            mutated = arch.copy()
            pos = int(rng.integers(lo, hi))
            mutated[pos] ^= 1 << int(rng.integers(0, 8))
            r = oracle.RefReader(mutated)
            rc_ref, out_ref = r.read(i)
            r.close()
            rc, out, _ = oracle.read_entry_port(2, mutated[lo:hi], size, size, int(d.hash[i]))
            assert (rc == 0) == (rc_ref == 0), (i, pos, rc, rc_ref)
            if rc_ref in (12, 13, 17):
                assert rc in (12, 13, 15, 17), (i, pos, rc, rc_ref)


# ------------------------------------------------------------------------------------------------ zstd
def test_zstd_frames_written_by_reference(oracle, zstd_cases):
    """Every zstd frame fixture written by the unmodified reference (levels 1..19: raw / RLE / compressed
    blocks, 1- and 4-stream Huffman literals, predefined / RLE / FSE / repeat sequence tables, multi-block
    frames) decodes bit-exactly with the restatement."""
    n = 0
    for k, comp in zstd_cases.items():
        if k.endswith("__in"):
            continue
        want = zstd_cases[k.split("__")[0] + "__in"]
        rc, out = oracle.zstd_decode_port(comp, len(want))
        assert rc == 0 and np.array_equal(out, want), k
        n += 1
    assert n >= 20


def test_zstd_error_classes_and_multiframe(oracle, zstd_cases):
    comp, want = zstd_cases["text_5k__l3"], zstd_cases["text_5k__in"]
    assert oracle.zstd_decode_port(comp[:-1], len(want))[0] == 13            # truncated
    assert oracle.zstd_decode_port(comp, len(want) - 1)[0] == 13             # dstSize_tooSmall is an error
    assert oracle.zstd_decode_port(np.concatenate([comp, np.zeros(2, np.uint8)]), len(want))[0] == 13  # trailing junk
    bad = comp.copy(); bad[0] ^= 1
    assert oracle.zstd_decode_port(bad, len(want))[0] == 13                  # magic
    skip = np.frombuffer(bytes([0x5A, 0x2A, 0x4D, 0x18, 3, 0, 0, 0]) + b"abc", np.uint8)
    two = np.concatenate([comp, skip, zstd_cases["one__l3"]])
    rc, out = oracle.zstd_decode_port(two, len(want) + 1)
    assert rc == 0 and np.array_equal(out, np.concatenate([want, zstd_cases["one__in"]]))
    rc, out = oracle.zstd_decode_port(np.zeros(0, np.uint8), 10)          # empty input: zero frames, zero bytes, no error
    assert rc == 0 and len(out) == 0


def test_ref_and_port_agree_on_zstd_corpus_and_corruption(oracle):
    """Differential run against the unmodified reference: its own compressor at several levels over the
    synthetic corpus (incl. sizes that give multi-block frames and the streamed-writer header shape via
    an unknown content size is not reachable through ZSTD_compress; covered by the golden archive), then
    bit flips / truncations: the entry-level verdict (OK vs not OK) must agree."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not present")
    from zpack_b200 import container, corpus
    rng = np.random.default_rng(11)
    for i, (size, lvl) in enumerate([(131072, 3), (131072, 3), (131072, 3), (131072, 3), (300000, 1), (70000, 5),
                                     (65536, 19), (1000, 3), (200000, 9), (13, 3)]):
        data = corpus.entry_bytes(i, size)
        comp = oracle.zstd_compress_ref(data, lvl)
        rc, out = oracle.zstd_decode_port(comp, size)
        assert rc == 0 and np.array_equal(out, data), (i, size, lvl)
        h = oracle.xxh3_port(data)
        for trial in range(12):
            m = comp.copy()
            if trial % 4 == 3:
                m = m[:int(rng.integers(1, len(m)))]
            else:
                m[int(rng.integers(0, len(m)))] ^= 1 << int(rng.integers(0, 8))
            arch = container.assemble(["x"], [m], [size], [h], [1])
            rd = oracle.RefReader(arch)
            rc_ref, _ = rd.read(0)
            rd.close()
            rc_port, _, _ = oracle.read_entry_port(1, m, size, size, h)
            assert (rc_ref == 0) == (rc_port == 0), (i, trial, rc_ref, rc_port)


def test_zstd_port_decodes_the_decodecorpus_frames(oracle, golden_dir):
    """Frames from zstd's own generator (externals/zstd/tests/decodecorpus.c, seeded; tests/golden/make_golden.py) and the
    upstream golden frame whose first block is RLE (tests/golden-decompression/rle-first-block.zst): sizes and XXH3-64
    digests recorded from the unmodified reference decoder."""
    d = dict(np.load(os.path.join(golden_dir, "zstd_corpus.npz")))
    names, sizes, dg = list(d["__names"]), d["__sizes"], d["__xxh3"]
    assert len(names) == 241 and "rle_first_block" in names
    for k, nm in enumerate(names):
        rc, got = oracle.zstd_decode_port(d[nm], int(sizes[k]))
        assert rc == 0 and len(got) == int(sizes[k]) and oracle.xxh3_port(got) == int(dg[k]), nm
        if oracle.have_ref() and k % 16 == 0:
            assert np.array_equal(oracle.zstd_decompress_ref(d[nm], int(sizes[k])), got), nm
