"""The container kernels (zpack_b200/csrc/archive_kernels.cuh: offset table, payload copy, central directory build and
parse) on the CPU emulation of tests/sim, compared with the host mirror (zpack_b200/container.py), the golden archives of
the reference's own tests and, when oracle/_ref is present, archives written by the unmodified reference.

The GPU parity tests of the same entry points are tests/test_gpu_archive.py (through the C-ABI on a B200)."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

from zpack_b200 import container
from zpack_b200.lib import ArcEntry

HERE = os.path.dirname(os.path.abspath(__file__))
SIM = os.path.join(HERE, "sim")
LIB = os.path.join(SIM, "libarchive_sim.so")
CSRC = os.path.join(os.path.dirname(HERE), "zpack_b200", "csrc")


def _build():
    deps = [os.path.join(SIM, f) for f in ("archive_sim.cpp", "sim_rt.h", "sim_cuda.h")]
    deps += [os.path.join(CSRC, f) for f in ("archive_kernels.cuh", "ptx.cuh", "common.cuh")]
    if os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-DZPB_SIM", "-shared", "-fPIC", "-w",
                    os.path.join(SIM, "archive_sim.cpp"), "-o", LIB], check=True)


@pytest.fixture(scope="module")
def sim():
    _build()
    lib = C.CDLL(LIB)
    vp, u64 = C.c_void_p, C.c_uint64
    lib.sim_archive_build.argtypes = [vp, vp, u64, vp, vp, C.c_int, u64, u64, C.c_int, C.c_int, u64, vp]
    lib.sim_archive_open.argtypes = [vp, u64, vp, u64, u64]
    lib.sim_archive_open.restype = u64
    return lib


def aligned(n, off=0):
    """n bytes starting `off` bytes past a 64-byte boundary, with slack on both sides (the copy reads whole 16-byte words)"""
    raw = np.zeros(n + 256, np.uint8)
    start = (-raw.ctypes.data) % 64 + 64 + off
    return raw[start:start + n], raw


def make_case(rng, sizes, name_lens, gap=7):
    """payloads in slots of a source buffer (arbitrary alignment), the entry table and the names blob"""
    n = len(sizes)
    payloads = [rng.integers(0, 256, s, dtype=np.uint8) for s in sizes]
    names = ["".join(chr(97 + int(c)) for c in rng.integers(0, 26, k)) for k in name_lens]
    e = np.zeros(n, ArcEntry)
    pos, npos = int(rng.integers(0, 16)), 0
    for i in range(n):
        e[i]["src_off"], e[i]["comp_size"], e[i]["uncomp_size"] = pos, sizes[i], sizes[i] * 3 + i
        e[i]["hash"], e[i]["method"] = int(rng.integers(0, 2**63)), i % 3
        e[i]["name_off"], e[i]["name_len"] = npos, name_lens[i]
        pos += sizes[i] + int(rng.integers(0, gap + 1))
        npos += name_lens[i]
    src, keep = aligned(pos + 16)
    for i in range(n):
        src[int(e[i]["src_off"]):int(e[i]["src_off"]) + sizes[i]] = payloads[i]
    blob = np.frombuffer("".join(names).encode(), np.uint8).copy() if npos else np.zeros(0, np.uint8)
    return payloads, names, e, src, keep, blob


def sim_build(sim, e, src, blob, seed=1, grid=3, scan_threads=128):
    n = len(e)
    data = int(e["comp_size"].sum())
    block = 35 * n + int(e["name_len"].sum())
    cdr_off = 10 + data
    total = cdr_off + 20 + block + 12
    out, keep = aligned(total, 0)
    totals = np.zeros(8, np.uint64)
    sim.sim_archive_build(src.ctypes.data, e.ctypes.data, n, blob.ctypes.data if len(blob) else None, out.ctypes.data, 1, cdr_off, block,
                          scan_threads, grid, seed, totals.ctypes.data)
    assert int(totals[0]) == data and int(totals[1]) == block
    return out


SIZES = [0, 1, 15, 16, 17, 31, 32, 33, 47, 48, 100, 4095, 65535, 65536, 65537, 131072 + 5, 200001, 0, 3]


def test_build_equals_the_host_mirror(sim):
    rng = np.random.default_rng(5)
    name_lens = [int(k) for k in rng.integers(1, 40, len(SIZES))]
    name_lens[3], name_lens[7] = 0, 700
    payloads, names, e, src, keep, blob = make_case(rng, SIZES, name_lens)
    ref = container.assemble(names, payloads, e["uncomp_size"], e["hash"], e["method"])
    for seed in range(1, 8):       # bits of the seed pick the copy variant in archive_sim.cpp: vectors in flight, deal, head grid
        out = sim_build(sim, e, src, blob, seed=seed)
        assert len(out) == len(ref) and np.array_equal(out, ref), seed
    assert np.array_equal(e["offset"], container.parse(ref).offset)


def test_build_many_small_entries_several_scan_rounds(sim):
    rng = np.random.default_rng(6)
    n = 700                                          # > 5 rounds of a 128-thread scan
    sizes = [int(s) for s in rng.integers(0, 90, n)]
    name_lens = [int(k) for k in rng.integers(0, 30, n)]
    payloads, names, e, src, keep, blob = make_case(rng, sizes, name_lens, gap=40)
    for seed in (1, 2):
        out = sim_build(sim, e.copy(), src, blob, seed=seed, grid=2)
        assert np.array_equal(out, container.assemble(names, payloads, e["uncomp_size"], e["hash"], e["method"]))


def test_copy_entries_to_given_offsets(sim):
    rng = np.random.default_rng(8)
    sizes = [70000, 5, 0, 16, 33000, 100]
    payloads, names, e, src, keep, blob = make_case(rng, sizes, [1] * len(sizes))
    order = [3, 0, 5, 1, 4, 2]
    pos = 3
    for i in order:
        e[i]["offset"] = pos
        pos += sizes[i] + 5
    dst, keep2 = aligned(pos + 8)
    dst[:] = 0xEE
    totals = np.zeros(8, np.uint64)
    sim.sim_archive_build(src.ctypes.data, e.ctypes.data, len(e), None, dst.ctypes.data, 0, 0, 0, 64, 2, 3, totals.ctypes.data)
    want = np.full(len(dst), 0xEE, np.uint8)
    for i in order:
        want[int(e[i]["offset"]):int(e[i]["offset"]) + sizes[i]] = payloads[i]
    assert np.array_equal(dst, want)                 # every byte outside the ranges untouched


def sim_open(sim, arch, count=None, block=None, seed=1):
    d = len(arch) - 12
    cdr_off = struct.unpack_from("<Q", arch, d + 4)[0]
    _, n, B = struct.unpack_from("<IQQ", arch, cdr_off)
    n = n if count is None else count
    B = B if block is None else block
    body = np.ascontiguousarray(arch[cdr_off + 20:cdr_off + 20 + B])
    out = np.zeros(max(n, 1), ArcEntry)
    found = sim.sim_archive_open(body.ctypes.data, B, out.ctypes.data, n, seed)
    return found, out[:n], body


def same_directory(out, body, d):
    assert np.array_equal(out["offset"], d.offset) and np.array_equal(out["src_off"], d.offset)
    assert np.array_equal(out["comp_size"], d.comp_size) and np.array_equal(out["uncomp_size"], d.uncomp_size)
    assert np.array_equal(out["hash"], d.hash) and np.array_equal(out["method"], d.method.astype(np.uint32))
    for i, nm in enumerate(d.names):
        o, k = int(out[i]["name_off"]), int(out[i]["name_len"])
        assert bytes(body[o:o + k]) == nm.encode("utf-8", "surrogateescape")


def test_open_equals_the_host_mirror(sim):
    rng = np.random.default_rng(11)
    n = 3000                                         # ~ 150 KB of directory: three super-tiles
    sizes = [int(s) for s in rng.integers(0, 50, n)]
    name_lens = [int(k) for k in rng.integers(0, 60, n)]
    name_lens[10], name_lens[1500], name_lens[1501] = 5000, 65535, 70    # records longer than a tile / a super-tile
    payloads, names, e, src, keep, blob = make_case(rng, sizes, name_lens)
    arch = container.assemble(names, payloads, e["uncomp_size"], e["hash"], e["method"])
    d = container.parse(arch)
    for seed in (1, 2):
        found, out, body = sim_open(sim, arch, seed=seed)
        assert found == n
        same_directory(out, body, d)
    # fewer files in the header than records in the block: the reference parses `count` records and ignores the rest
    found, out, body = sim_open(sim, arch, count=n - 7)
    assert found >= n - 7
    assert np.array_equal(out["hash"], d.hash[:n - 7])
    # a block cut inside a record: the walk finds fewer records than the header promises (BLOCK_SIZE_INVALID in the C-ABI)
    _, _, B = struct.unpack_from("<IQQ", arch, d.cdr_offset)
    found, _, _ = sim_open(sim, arch, block=B - 1)
    assert found == n - 1
    found, _, _ = sim_open(sim, arch, block=34)
    assert found == 0


def test_open_golden_and_reference_written_archives(sim, golden_dir, oracle):
    archives = [np.fromfile(os.path.join(golden_dir, f), np.uint8) for f in sorted(os.listdir(golden_dir)) if f.endswith(".zpk")]
    assert archives
    if oracle.have_ref():
        bufs = [np.frombuffer(bytes([i % 251]) * (100 + 37 * i), np.uint8) for i in range(200)]
        archives.append(oracle.write_archive_ref([f"dir{i % 7}/file_{i:05d}.bin" for i in range(200)], bufs, 2, 0))
    for arch in archives:
        d = container.parse(arch)
        found, out, body = sim_open(sim, arch)
        assert found == len(d)
        same_directory(out, body, d)


def test_copy_all_alignments_and_short_lengths(sim):
    """every (source, destination) phase mod 16 x lengths around the vector / head / tail boundaries, all copy variants"""
    rng = np.random.default_rng(31)
    lens = [0, 1, 2, 15, 16, 17, 31, 32, 33, 47, 48, 63, 64, 65, 127, 128, 129, 255, 511, 2047, 2049, 4097]
    cases = [(sp, dp, n) for sp in range(16) for dp in range(16) for n in (lens[(sp * 16 + dp + k) % len(lens)] for k in range(2))]
    e = np.zeros(len(cases), ArcEntry)
    spos = dpos = 0
    for i, (sp, dp, n) in enumerate(cases):
        spos += (sp - spos) % 16
        dpos += (dp - dpos) % 16
        e[i]["src_off"], e[i]["offset"], e[i]["comp_size"] = spos, dpos, n
        spos += n + int(rng.integers(0, 9))
        dpos += n + int(rng.integers(0, 9))
    src, k1 = aligned(spos + 16)
    src[:] = rng.integers(0, 256, len(src), dtype=np.uint8)
    want = np.full(dpos + 16, 0xEE, np.uint8)
    for r in e:
        want[int(r["offset"]):int(r["offset"]) + int(r["comp_size"])] = src[int(r["src_off"]):int(r["src_off"]) + int(r["comp_size"])]
    for seed in range(1, 8):
        dst, k2 = aligned(dpos + 16)
        dst[:] = 0xEE
        totals = np.zeros(8, np.uint64)
        sim.sim_archive_build(src.ctypes.data, e.ctypes.data, len(e), None, dst.ctypes.data, 0, 0, 0, 64, 3, seed, totals.ctypes.data)
        assert np.array_equal(dst, want), seed


def test_open_random_directories(sim):
    """random record shapes (empty names, names longer than a tile, directories ending exactly on a tile edge) against the host mirror"""
    rng = np.random.default_rng(32)
    for trial in range(12):
        n = int(rng.integers(1, 500))
        kind = trial % 4
        if kind == 0:
            name_lens = [0] * n
        elif kind == 1:
            name_lens = [int(k) for k in rng.integers(0, 200, n)]
        elif kind == 2:
            name_lens = [int(k) if rng.random() > 0.03 else int(rng.integers(1000, 9000)) for k in rng.integers(0, 30, n)]
        else:                       # 35 + 93 = 128 bytes per record: records end on every tile edge
            name_lens = [93] * n
        payloads, names, e, src, keep, blob = make_case(rng, [0] * n, name_lens)
        arch = container.assemble(names, payloads, e["uncomp_size"], e["hash"], e["method"])
        d = container.parse(arch)
        found, out, body = sim_open(sim, arch, seed=trial + 1)
        assert found == n, (trial, found, n)
        same_directory(out, body, d)
        d2 = container.directory_from_table(out, body, len(arch))      # the Python view of a device-side open
        assert d2.names == d.names and d2.cdr_offset == d.cdr_offset and d2.file_size == d.file_size
        assert np.array_equal(d2.entries(), d.entries())


def test_build_is_byte_identical_to_the_reference_writer(sim, oracle):
    """Stored (method NONE) files: the archive the kernels assemble from the raw files + their XXH3-64 digests is, byte for byte,
    the archive the unmodified reference's zpack_write_archive produces for the same files."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not present")
    rng = np.random.default_rng(41)
    sizes = [0, 1, 17, 4096, 70001, 5, 131072, 33]
    names = [f"dir{i % 3}/file_{i:03d}.dat" for i in range(len(sizes))]
    bufs = [rng.integers(0, 256, s, dtype=np.uint8) for s in sizes]
    ref = oracle.write_archive_ref(names, bufs, 0, 0)
    e = np.zeros(len(sizes), ArcEntry)
    pos = npos = 0
    for i, b in enumerate(bufs):
        e[i]["src_off"], e[i]["comp_size"], e[i]["uncomp_size"], e[i]["hash"], e[i]["method"] = pos, len(b), len(b), oracle.xxh3_port(b), 0
        e[i]["name_off"], e[i]["name_len"] = npos, len(names[i])
        pos += len(b) + 11
        npos += len(names[i])
    src, keep = aligned(pos + 16)
    for r, b in zip(e, bufs):
        src[int(r["src_off"]):int(r["src_off"]) + len(b)] = b
    blob = np.frombuffer("".join(names).encode(), np.uint8).copy()
    for seed in (1, 6):
        out = sim_build(sim, e.copy(), src, blob, seed=seed)
        assert len(out) == len(ref) and np.array_equal(out, ref), seed
