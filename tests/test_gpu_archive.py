"""The container level on the device through the C-ABI (zpb_archive_open_device / zpb_archive_build_device /
zpb_copy_entries_device): results equal the host mirror (zpack_b200/container.py), the reference's golden archives parse to
the same directory, and the unmodified reference reader accepts an archive that was packed, assembled and given its central
directory without leaving HBM."""
import os
import struct

import numpy as np
import pytest

from zpack_b200 import container, corpus
from zpack_b200.lib import ArcEntry, File

pytestmark = pytest.mark.gpu


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def table(sizes, name_lens, rng, gap=7):
    n = len(sizes)
    e = np.zeros(n, ArcEntry)
    names = ["".join(chr(97 + int(c)) for c in rng.integers(0, 26, k)) for k in name_lens]
    pos, npos = int(rng.integers(0, 16)), 0
    for i in range(n):
        e[i]["src_off"], e[i]["comp_size"], e[i]["uncomp_size"] = pos, sizes[i], sizes[i] * 2 + 1
        e[i]["hash"], e[i]["method"] = int(rng.integers(0, 2**63)), i % 3
        e[i]["name_off"], e[i]["name_len"] = npos, name_lens[i]
        pos += sizes[i] + int(rng.integers(0, gap + 1))
        npos += name_lens[i]
    src = rng.integers(0, 256, pos + 16, dtype=np.uint8)
    blob = np.frombuffer("".join(names).encode(), np.uint8).copy()
    payloads = [src[int(e[i]["src_off"]):int(e[i]["src_off"]) + sizes[i]] for i in range(n)]
    return e, names, blob, src, payloads


def same_directory(e, names_blob, d):
    assert np.array_equal(e["offset"], d.offset) and np.array_equal(e["comp_size"], d.comp_size)
    assert np.array_equal(e["uncomp_size"], d.uncomp_size) and np.array_equal(e["hash"], d.hash)
    assert np.array_equal(e["method"], d.method.astype(np.uint32))
    for i, nm in enumerate(d.names):
        o, k = int(e[i]["name_off"]), int(e[i]["name_len"])
        assert bytes(names_blob[o:o + k]) == nm.encode("utf-8", "surrogateescape")


def test_build_then_open_equal_the_host_mirror(gpu_ctx):
    import torch
    rng = np.random.default_rng(21)
    sizes = [0, 1, 15, 16, 17, 33, 4095, 65535, 65536, 65537, 200001, 0, 3] + [int(s) for s in rng.integers(0, 150000, 400)]
    name_lens = [int(k) for k in rng.integers(0, 48, len(sizes))]
    name_lens[5], name_lens[100], name_lens[101] = 5000, 65535, 1
    e, names, blob, src, payloads = table(sizes, name_lens, rng)
    ref = container.assemble(names, payloads, e["uncomp_size"], e["hash"], e["method"])
    d_src = dev(src)
    d_arch = torch.full((len(ref) + 64,), 0xEE, dtype=torch.uint8, device="cuda")
    size = gpu_ctx.archive_build_device(d_src, len(src), e, blob, d_arch, len(ref) + 64)
    assert size == len(ref)
    got = d_arch.cpu().numpy()
    assert np.array_equal(got[:size], ref) and (got[size:] == 0xEE).all()
    d = container.parse(ref)
    assert np.array_equal(e["offset"], d.offset)                          # the offset table came back
    res, e2, nb = gpu_ctx.archive_open_device(d_arch, size)
    assert res == 0 and len(e2) == len(d)
    same_directory(e2, nb, d)
    with pytest.raises(Exception):                                        # too small an archive buffer is refused, not overrun
        gpu_ctx.archive_build_device(d_src, len(src), e, blob, d_arch, size - 1)


def test_open_large_directory_and_the_reference_results(gpu_ctx, oracle, golden_dir):
    import torch
    n = 70000                                                             # ~ 4 MB of directory, 60 super-tiles
    rng = np.random.default_rng(22)
    name_lens = rng.integers(8, 40, n)
    names = [f"d{i % 97}/" + "x" * (int(k) - 6) + f"{i:06d}"[:6] for i, k in enumerate(name_lens)]
    comp = rng.integers(0, 1000, n).astype(np.uint64)
    cdr = container.cdr_bytes(names, 10 + np.concatenate([[0], np.cumsum(comp)[:-1]]).astype(np.uint64), comp, comp * 3,
                              rng.integers(0, 2**63, n).astype(np.uint64), rng.integers(0, 3, n))
    cdr_off = 10 + int(comp.sum())
    arch = np.zeros(cdr_off + len(cdr) + 12, np.uint8)
    arch[:10] = np.frombuffer(struct.pack("<IHI", container.SIG_HEADER, 1, container.SIG_DATA), np.uint8)
    arch[cdr_off:cdr_off + len(cdr)] = np.frombuffer(cdr, np.uint8)
    arch[-12:] = np.frombuffer(struct.pack("<IQ", container.SIG_EOCDR, cdr_off), np.uint8)
    d = container.parse(arch)
    res, e, nb = gpu_ctx.archive_open_device(dev(arch), len(arch))
    assert res == 0
    same_directory(e, nb, d)
    print(f"\nCDR parse of {n} entries ({len(cdr) / 1e6:.1f} MB): {gpu_ctx.last_archive_ms()[2]:.3f} ms of kernels")
    archives = [np.fromfile(os.path.join(golden_dir, f), np.uint8) for f in sorted(os.listdir(golden_dir)) if f.endswith(".zpk")]
    if oracle.have_ref():
        bufs = [corpus.entry_bytes(i, 3000 + 11 * i) for i in range(300)]
        archives.append(oracle.write_archive_ref([corpus.entry_name(i) for i in range(300)], bufs, 2, 0))
    for a in archives:
        res, e, nb = gpu_ctx.archive_open_device(dev(a), len(a))
        assert res == 0
        same_directory(e, nb, container.parse(a))


def test_open_refuses_what_the_reference_refuses(gpu_ctx, oracle):
    rng = np.random.default_rng(23)
    e, names, blob, src, payloads = table([100, 2000, 30], [4, 9, 300], rng)
    good = container.assemble(names, payloads, e["uncomp_size"], e["hash"], e["method"])
    cdr_off = container.parse(good).cdr_offset

    def opened(a):
        return gpu_ctx.archive_open_device(dev(a), len(a))[0]

    def mutated(fn):
        a = good.copy()
        fn(a)
        return a

    cases = {
        "short": good[:41],
        "header signature": mutated(lambda a: a.__setitem__(0, 0)),
        "version": mutated(lambda a: a.__setitem__(4, 2)),
        "data signature": mutated(lambda a: a.__setitem__(6, 0)),
        "eocdr signature": mutated(lambda a: a.__setitem__(len(a) - 12, 0)),
        "cdr offset past the end": mutated(lambda a: a.__setitem__(slice(len(a) - 8, len(a)), np.frombuffer(struct.pack("<Q", len(a)), np.uint8))),
        "cdr signature": mutated(lambda a: a.__setitem__(cdr_off, 0)),
        "block size too large": mutated(lambda a: a.__setitem__(slice(cdr_off + 12, cdr_off + 20), np.frombuffer(struct.pack("<Q", 10**6), np.uint8))),
        "count too large": mutated(lambda a: a.__setitem__(slice(cdr_off + 4, cdr_off + 12), np.frombuffer(struct.pack("<Q", 4), np.uint8))),
        "count beyond the fixed sizes": mutated(lambda a: a.__setitem__(slice(cdr_off + 4, cdr_off + 12), np.frombuffer(struct.pack("<Q", 10**9), np.uint8))),
        "name length past the block": mutated(lambda a: a.__setitem__(slice(cdr_off + 20, cdr_off + 22), np.frombuffer(struct.pack("<H", 60000), np.uint8))),
    }
    want = {"short": 5, "header signature": 6, "version": 9, "data signature": 6, "eocdr signature": 6, "cdr offset past the end": 7,
            "cdr signature": 6, "block size too large": 8, "count too large": 8, "count beyond the fixed sizes": 8,
            "name length past the block": 8}
    for k, a in cases.items():
        got = opened(a)
        assert got == want[k], (k, got)
        if oracle.have_ref():
            assert oracle.RefReader.open_result(a) == got, k
    assert opened(good) == 0


def test_copy_entries_between_archives_on_the_device(gpu_ctx):
    """zpack_write_files_from_archive with both archives in HBM: a subset of one archive's entries becomes a new archive."""
    import torch
    rng = np.random.default_rng(24)
    sizes = [int(s) for s in rng.integers(0, 300000, 200)]
    e, names, blob, src, payloads = table(sizes, [int(k) for k in rng.integers(1, 30, 200)], rng)
    old = container.assemble(names, payloads, e["uncomp_size"], e["hash"], e["method"])
    d_old = dev(old)
    res, eo, nb = gpu_ctx.archive_open_device(d_old, len(old))
    assert res == 0
    keep = [i for i in range(200) if i % 3 != 1][::-1]                                # delete a third, reverse the order
    sub = np.ascontiguousarray(eo[keep])                                              # src_off = the old archive's offsets
    want = container.assemble([names[i] for i in keep], [payloads[i] for i in keep], e["uncomp_size"][keep], e["hash"][keep], e["method"][keep])
    d_new = torch.zeros(len(want) + 16, dtype=torch.uint8, device="cuda")
    size = gpu_ctx.archive_build_device(d_old, len(old), sub, nb, d_new, len(want) + 16)   # names: the old directory block
    assert size == len(want) and np.array_equal(d_new.cpu().numpy()[:size], want)
    # the bare copy, destination offsets chosen by the caller
    sub2 = sub.copy()
    pos = 5
    for r in sub2:
        r["offset"] = pos
        pos += int(r["comp_size"]) + 3
    d_dst = torch.full((pos + 32,), 0x77, dtype=torch.uint8, device="cuda")
    gpu_ctx.copy_entries_device(d_old, len(old), d_dst, pos + 32, sub2)
    got = d_dst.cpu().numpy()
    exp = np.full(pos + 32, 0x77, np.uint8)
    for r, i in zip(sub2, keep):
        exp[int(r["offset"]):int(r["offset"]) + sizes[i]] = payloads[i]
    assert np.array_equal(got, exp)
    sub2[0]["offset"] = pos + 100
    with pytest.raises(Exception):
        gpu_ctx.copy_entries_device(d_old, len(old), d_dst, pos + 32, sub2)


def test_pack_assemble_and_read_back_without_leaving_the_device(gpu_ctx, oracle):
    """pack -> archive image -> open -> unpack, every step on device-resident buffers; the reference reader agrees."""
    import torch
    n, size = 96, 131072
    bufs = [corpus.entry_bytes(i, size if i % 5 else size // 7) for i in range(n)]
    files = np.zeros(n, File)
    pos = opos = 0
    for i, b in enumerate(bufs):
        cap = gpu_ctx.pack_bound(2, len(b))
        files[i] = (pos, len(b), opos, cap, 2, 0, (0, 0, 0))
        pos += (len(b) + 15) & ~15
        opos += (cap + 15) & ~15
    h_in = np.zeros(pos, np.uint8)
    for f, b in zip(files, bufs):
        h_in[int(f["src_off"]):int(f["src_off"]) + len(b)] = b
    d_in, d_slots = dev(h_in), torch.zeros(opos, dtype=torch.uint8, device="cuda")
    comp, digest, status = gpu_ctx.pack_device(d_in, pos, d_slots, opos, files)
    assert (status == 0).all()
    names = [corpus.entry_name(i) for i in range(n)]
    e = np.zeros(n, ArcEntry)
    e["src_off"], e["comp_size"], e["uncomp_size"], e["hash"], e["method"] = files["dst_off"], comp, files["size"], digest, 2
    e["name_len"] = [len(s) for s in names]
    e["name_off"] = np.concatenate([[0], np.cumsum(e["name_len"])[:-1]])
    blob = np.frombuffer("".join(names).encode(), np.uint8)
    cap = 10 + int(comp.sum()) + 20 + 35 * n + len(blob) + 12
    d_arch = torch.zeros(cap, dtype=torch.uint8, device="cuda")
    asize = gpu_ctx.archive_build_device(d_slots, opos, e, blob, d_arch, cap)
    assert asize == cap
    res, eo, nb = gpu_ctx.archive_open_device(d_arch, asize)
    assert res == 0 and np.array_equal(eo["hash"], digest)
    arch = d_arch.cpu().numpy()
    ent = container.parse(arch).entries()
    out_size = int((ent["dst_off"] + ent["dst_cap"]).max())
    d_out = torch.zeros(out_size + 16, dtype=torch.uint8, device="cuda")
    st, dg = gpu_ctx.unpack_device(d_arch, asize, d_out, out_size, ent)
    assert (st == 0).all() and np.array_equal(dg, digest)
    out = d_out.cpu().numpy()
    for i, b in enumerate(bufs):
        o = int(ent["dst_off"][i])
        assert np.array_equal(out[o:o + len(b)], b)
    if oracle.have_ref():
        rd = oracle.RefReader(arch)
        assert rd.count == n
        for i in (0, 1, n // 2, n - 1):
            rc, got = rd.read(i)
            assert rc == 0 and np.array_equal(got, bufs[i])


def test_file_to_device_and_back(gpu_ctx, tmp_path):
    """zpb_file_read_device / zpb_file_write_device: ranges of a file <-> HBM through the pinned pipeline (several lanes,
    more than two rounds per lane, a ragged tail), and the error for a file that is too short."""
    import torch
    rng = np.random.default_rng(25)
    size = (72 << 20) + 12345
    data = rng.integers(0, 256, size, dtype=np.uint8)
    p = tmp_path / "blob.bin"
    data.tofile(p)
    fd = os.open(p, os.O_RDONLY)
    try:
        for off, n in ((0, size), (7, 100), (size - 5, 5), (3, (9 << 20) + 1), (11, 0)):
            d = torch.zeros(n + 16, dtype=torch.uint8, device="cuda")
            gpu_ctx.file_read_device(fd, off, n, d)
            got = d.cpu().numpy()
            assert np.array_equal(got[:n], data[off:off + n]) and (got[n:] == 0).all(), (off, n)
        with pytest.raises(Exception):
            gpu_ctx.file_read_device(fd, size - 10, 100, torch.zeros(128, dtype=torch.uint8, device="cuda"))
    finally:
        os.close(fd)
    q = tmp_path / "out.bin"
    fd = os.open(q, os.O_RDWR | os.O_CREAT)
    try:
        d = dev(data)
        gpu_ctx.file_write_device(fd, 5, size, d)
        gpu_ctx.file_write_device(fd, 0, 5, d)
    finally:
        os.close(fd)
    back = np.fromfile(q, np.uint8)
    assert len(back) == size + 5 and np.array_equal(back[5:], data) and np.array_equal(back[:5], data[:5])


def test_archive_file_to_verified_plaintext_without_a_host_copy_of_the_archive(gpu_ctx, oracle, tmp_path):
    """file -> HBM -> directory parse -> unpack + verify: the host sees the 42 fixed bytes, the entry table and the names."""
    import torch
    n = 64
    bufs = [corpus.entry_bytes(i, 100000 + 997 * i) for i in range(n)]
    frames = [oracle.lz4f_encode_port(b, 0) for b in bufs]
    hashes = [oracle.xxh3_port(b) for b in bufs]
    arch = container.assemble([corpus.entry_name(i) for i in range(n)], frames, [len(b) for b in bufs], hashes, [2] * n)
    p = tmp_path / "a.zpk"
    arch.tofile(p)
    fd = os.open(p, os.O_RDONLY)
    try:
        d_arch = torch.zeros(len(arch) + 16, dtype=torch.uint8, device="cuda")
        gpu_ctx.file_read_device(fd, 0, len(arch), d_arch)
    finally:
        os.close(fd)
    res, e, nb = gpu_ctx.archive_open_device(d_arch, len(arch))
    assert res == 0 and len(e) == n
    ent = np.zeros(n, zpack_entry_dtype())
    ent["src_off"], ent["comp_size"], ent["uncomp_size"], ent["hash"], ent["method"] = e["offset"], e["comp_size"], e["uncomp_size"], e["hash"], e["method"]
    ent["dst_cap"] = e["uncomp_size"]
    padded = (e["uncomp_size"] + np.uint64(15)) & ~np.uint64(15)
    ent["dst_off"] = np.concatenate([[0], np.cumsum(padded)[:-1]])
    out_size = int(padded.sum())
    d_out = torch.zeros(out_size + 16, dtype=torch.uint8, device="cuda")
    st, dg = gpu_ctx.unpack_device(d_arch, len(arch), d_out, out_size, ent)
    assert (st == 0).all() and np.array_equal(dg, np.array(hashes, np.uint64))
    out = d_out.cpu().numpy()
    for i, b in enumerate(bufs):
        o = int(ent["dst_off"][i])
        assert np.array_equal(out[o:o + len(b)], b)


def zpack_entry_dtype():
    from zpack_b200.lib import Entry
    return Entry


def test_empty_archive_and_empty_entries(gpu_ctx, oracle):
    """n = 0 (the 42-byte archive the reference writes for no files) and entries of size 0 only."""
    import torch
    ref0 = container.assemble([], [], [], [], [])
    assert len(ref0) == 42
    d_arch = torch.full((64,), 0xEE, dtype=torch.uint8, device="cuda")
    size = gpu_ctx.archive_build_device(None, 0, np.zeros(0, ArcEntry), np.zeros(0, np.uint8), d_arch, 64)
    assert size == 42 and np.array_equal(d_arch.cpu().numpy()[:42], ref0)
    res, e, nb = gpu_ctx.archive_open_device(d_arch, 42)
    assert res == 0 and len(e) == 0
    if oracle.have_ref():
        assert oracle.RefReader.open_result(ref0) == 0
    names = ["a", "", "ccc"]
    e = np.zeros(3, ArcEntry)
    e["name_len"], e["name_off"], e["method"], e["hash"] = [1, 0, 3], [0, 1, 1], [0, 2, 1], [7, 8, 9]
    blob = np.frombuffer(b"accc", np.uint8)
    ref = container.assemble(names, [b"", b"", b""], [0, 0, 0], [7, 8, 9], [0, 2, 1])
    d_arch = torch.zeros(len(ref) + 16, dtype=torch.uint8, device="cuda")
    size = gpu_ctx.archive_build_device(None, 0, e, blob, d_arch, len(ref) + 16)
    assert size == len(ref) and np.array_equal(d_arch.cpu().numpy()[:size], ref)
    assert (e["offset"] == 10).all()
    res, eo, nb = gpu_ctx.archive_open_device(d_arch, size)
    assert res == 0 and list(eo["name_len"]) == [1, 0, 3] and list(eo["hash"]) == [7, 8, 9]
    gpu_ctx.copy_entries_device(d_arch, size, d_arch, size, np.zeros(0, ArcEntry))      # nothing to do is not an error


def test_build_is_byte_identical_to_the_reference_writer_and_open_equals_its_reader(gpu_ctx, oracle):
    """Stored files + their XXH3-64 digests (computed on the GPU): zpb_archive_build_device writes, byte for byte, the archive
    the unmodified reference's zpack_write_archive writes; zpb_archive_open_device returns the unmodified reader's entries."""
    import torch
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not present")
    rng = np.random.default_rng(41)
    sizes = [0, 1, 17, 4096, 70001, 5, 131072, 33, 300000]
    names = [f"dir{i % 3}/file_{i:03d}.dat" for i in range(len(sizes))]
    bufs = [rng.integers(0, 256, s, dtype=np.uint8) for s in sizes]
    ref = oracle.write_archive_ref(names, bufs, 0, 0)
    e = np.zeros(len(sizes), ArcEntry)
    pos = npos = 0
    for i, b in enumerate(bufs):
        e[i]["src_off"], e[i]["comp_size"], e[i]["uncomp_size"], e[i]["method"] = pos, len(b), len(b), 0
        e[i]["name_off"], e[i]["name_len"] = npos, len(names[i])
        pos += (len(b) + 15) & ~15
        npos += len(names[i])
    src = np.zeros(pos + 16, np.uint8)
    for r, b in zip(e, bufs):
        src[int(r["src_off"]):int(r["src_off"]) + len(b)] = b
    d_src = dev(src)
    e["hash"] = gpu_ctx.xxh3_device(d_src, e["src_off"], e["comp_size"])
    blob = np.frombuffer("".join(names).encode(), np.uint8)
    d_arch = torch.zeros(len(ref) + 16, dtype=torch.uint8, device="cuda")
    size = gpu_ctx.archive_build_device(d_src, len(src), e, blob, d_arch, len(ref) + 16)
    assert size == len(ref) and np.array_equal(d_arch.cpu().numpy()[:size], ref)
    res, eo, nb = gpu_ctx.archive_open_device(dev(ref), len(ref))
    assert res == 0
    rd = oracle.RefReader(ref)
    assert rd.count == len(eo)
    for i in range(rd.count):
        r = rd.entry(i)
        o, k = int(eo[i]["name_off"]), int(eo[i]["name_len"])
        assert (int(r.offset), int(r.comp_size), int(r.uncomp_size), int(r.hash), int(r.comp_method)) == \
               (int(eo[i]["offset"]), int(eo[i]["comp_size"]), int(eo[i]["uncomp_size"]), int(eo[i]["hash"]), int(eo[i]["method"]))
        assert r.filename == bytes(nb[o:o + k])
    rd.close()
