"""The LZ4 scan / parse / execute kernels (zpack_b200/csrc/lz4_fast.cuh) run on a CPU emulation of the CUDA execution
model (tests/sim/sim_rt.h: fibers per thread, warp collectives, random lane order between synchronisation points,
shared-memory bounds and race checks, deferred completion of cp.async / cp.async.bulk) and are compared with the oracle.

This does not replace the GPU parity tests (tests/test_gpu_*.py call the real library through its C-ABI on a B200);
it lets the kernels' control logic — sequence field derivation, ring addressing, staging rows and barrier phases,
the same-step dependency pass, lane / warp copy selection — be checked in a container that has no GPU, on the same
kernel source that nvcc compiles (g++ -DZPB_SIM swaps the PTX wrappers of csrc/ptx.cuh for emulated ones)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from zpack_b200 import container, corpus
from zpack_b200.lib import Entry

import crafted

HERE = os.path.dirname(os.path.abspath(__file__))
SIM = os.path.join(HERE, "sim")
LIB = os.path.join(SIM, "liblz4_sim.so")
CSRC = os.path.join(os.path.dirname(HERE), "zpack_b200", "csrc")


def _build():
    deps = [os.path.join(SIM, f) for f in ("lz4_sim.cpp", "sim_rt.h", "sim_cuda.h")]
    deps += [os.path.join(CSRC, f) for f in ("lz4_fast.cuh", "ptx.cuh", "common.cuh", "xxh3.cuh", "lz4_decode.cuh")]
    if os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-DZPB_SIM", "-shared", "-fPIC", "-w",
                    os.path.join(SIM, "lz4_sim.cpp"), "-o", LIB], check=True)


@pytest.fixture(scope="module")
def sim():
    _build()
    lib = C.CDLL(LIB)
    vp = C.c_void_p
    lib.sim_lz4_unpack.argtypes = [vp, C.c_uint64, vp, C.c_uint64, vp, C.c_uint64, vp, vp, C.c_int, C.c_int,
                                   C.c_uint64, vp, vp]
    return lib


def run(sim, arch, entries, out_size, seed=1, grid_parse=1, grid_exec=1):
    n = len(entries)
    out = np.zeros(out_size + 64, np.uint8)
    status = np.zeros(n, np.int32)
    digest = np.zeros(n, np.uint64)
    races = np.zeros(1, np.uint64)
    ngen = sim.sim_lz4_unpack(arch.ctypes.data, len(arch), out.ctypes.data, out_size, entries.ctypes.data, n,
                              status.ctypes.data, digest.ctypes.data, grid_parse, grid_exec, seed, races.ctypes.data, None)
    return out, status, digest, int(races[0]), ngen


def check(sim, oracle, frames, plain, methods=None, seed=1, expect_general=0):
    hashes = [oracle.xxh3_port(b) for b in plain]
    methods = methods or [2] * len(frames)
    arch = container.assemble([f"e{i}" for i in range(len(frames))], frames, [len(b) for b in plain], hashes, methods)
    # the archive does not start on a 16-byte boundary inside its buffer: exercises the staging skew and edge bytes
    buf = np.zeros(len(arch) + 64, np.uint8)
    off = 5
    base = (-buf.ctypes.data) % 16 + off
    buf[base:base + len(arch)] = arch
    view = buf[base:base + len(arch)]
    e = container.parse(arch).entries()
    out_size = int((e["dst_off"] + e["dst_cap"]).max()) if len(e) else 0
    n = len(e)
    out = np.zeros(out_size + 80, np.uint8)
    obase = (-out.ctypes.data) % 16
    status = np.zeros(n, np.int32)
    digest = np.zeros(n, np.uint64)
    races = np.zeros(1, np.uint64)
    ngen = sim.sim_lz4_unpack(view.ctypes.data, len(arch), out.ctypes.data + obase, out_size, e.ctypes.data, n,
                              status.ctypes.data, digest.ctypes.data, 1, 1, seed, races.ctypes.data, None)
    assert int(races[0]) == 0, "shared-memory races reported by the emulation"
    assert ngen == expect_general
    if expect_general == 0:
        assert (status == 0).all(), status
        hashed = e["comp_size"] != 0   # comp_size == 0: OK without hashing (lib/zpack_read.c:328)
        assert np.array_equal(digest[hashed], np.array(hashes, np.uint64)[hashed])
    for i, b in enumerate(plain):
        if status[i] != 0:
            continue
        o = obase + int(e["dst_off"][i])
        got = out[o:o + len(b)]
        if not np.array_equal(got, b):
            bad = int(np.argmax(got != b))
            raise AssertionError(f"entry {i}: first wrong byte at {bad} of {len(b)}")
    return status


def test_corpus_entries_reference_frames(sim, oracle):
    """The bench corpus' four classes as the reference writes them (block-linked 64 KB blocks)."""
    frames, plain = [], []
    for i in range(8):
        b = corpus.entry_bytes(i, 131072 if i < 4 else 70001)
        frames.append(oracle.lz4f_encode_port(b, 0, independent=False))
        plain.append(b)
    check(sim, oracle, frames, plain)


def test_corpus_entries_independent_blocks_and_raw(sim, oracle):
    frames, plain, methods = [], [], []
    for i in range(4):
        b = corpus.entry_bytes(i, 65536 + 4096)
        frames.append(oracle.lz4f_encode_port(b, 0, independent=True))
        plain.append(b)
        methods.append(2)
    for size in (0, 1, 15, 16, 17, 240, 241, 1024, 1025, 5000):
        b = corpus.entry_bytes(1, size)
        frames.append(b.copy() if size else np.zeros(0, np.uint8))   # ZPACK_COMPRESSION_NONE
        plain.append(b)
        methods.append(0)
        frames.append(oracle.lz4f_encode_port(b, 0, independent=False))
        plain.append(b)
        methods.append(2)
    check(sim, oracle, frames, plain, methods)


@pytest.mark.parametrize("style", ["mixed", "chains", "fills", "long"])
def test_crafted_frames(sim, oracle, style):
    rng = np.random.default_rng(0xC0FFEE + hash(style) % 1000)
    frames, plain = [], []
    for sizes in ([65536], [65536, 30000], [20000], [65536, 1000]):
        fr, data = crafted.frame(rng, sizes, style)
        rc, got = oracle.lz4f_decode_port(np.frombuffer(fr, np.uint8), len(data))
        assert rc == 0 and bytes(got) == data, "builder and oracle disagree"
        frames.append(np.frombuffer(fr, np.uint8))
        plain.append(np.frombuffer(data, np.uint8))
    check(sim, oracle, frames, plain, seed=7)


def test_golden_lz4_cases(sim, oracle, lz4_cases):
    """Frames written by the unmodified reference (tests/golden/make_golden.py); shapes the fast path declines
    (checksums, > 64 KB blocks, multi-frame) must be handed to the general decoder, never decoded wrongly."""
    names = sorted(k for k in lz4_cases if not k.endswith("__in"))
    frames = [lz4_cases[k] for k in names]
    plain = [lz4_cases[k.split("__")[0] + "__in"] for k in names]
    hashes = [oracle.xxh3_port(b) for b in plain]
    arch = container.assemble(names, frames, [len(b) for b in plain], hashes, [2] * len(frames))
    e = container.parse(arch).entries()
    out_size = int((e["dst_off"] + e["dst_cap"]).max())
    out, status, digest, races, ngen = run(sim, arch, e, out_size)
    assert races == 0
    assert (status == 0).sum() + ngen == len(names)
    for i, b in enumerate(plain):
        if status[i] == 0:
            o = int(e["dst_off"][i])
            assert np.array_equal(out[o:o + len(b)], b), names[i]
            assert int(digest[i]) == hashes[i]
        else:
            assert status[i] == -1000, (names[i], status[i])   # declined, not mis-decoded
    declined = {names[i] for i in range(len(names)) if status[i] != 0}
    assert all(any(t in k for t in ("sums", "256k", "4m")) for k in declined), declined


def test_split_parse_on_sparse_matches(sim, oracle):
    """Blocks the four-way split parse cannot join (mostly literals, a match every few hundred bytes: a walk that starts
    inside a literal run stays off the true chain) must be walked again unsplit and still decode."""
    rng = np.random.default_rng(11)
    frames, plain = [], []
    for k in range(4):
        base = rng.integers(0, 256, size=131072, dtype=np.uint8)
        for _ in range(150 * (k + 1)):                       # sprinkle short repeats
            p = int(rng.integers(100, 131072 - 40)); q = int(rng.integers(0, p - 20)); n = int(rng.integers(6, 30))
            base[p:p + n] = base[q:q + n]
        frames.append(oracle.lz4f_encode_port(base, 0, independent=False))
        plain.append(base)
    check(sim, oracle, frames, plain, seed=3)


def test_corrupted_frames_are_declined_not_misdecoded(sim, oracle):
    """Bit flips inside the blocks of reference-format frames: the emulation's bounds checks must stay silent (memory
    safety of the speculative walks on garbage), and every entry must end as decoded-and-verified, hash mismatch, or
    handed to the general decoder — never as OK with wrong bytes."""
    rng = np.random.default_rng(5)
    frames, plain = [], []
    for i in range(12):
        b = corpus.entry_bytes(1 + 2 * (i % 2), 131072)      # text and records
        f = oracle.lz4f_encode_port(b, 0, independent=False).copy()
        for _ in range(1 + i % 3):
            p = int(rng.integers(11, len(f) - 4))
            f[p] ^= 1 << int(rng.integers(0, 8))
        frames.append(f)
        plain.append(b)
    hashes = [oracle.xxh3_port(b) for b in plain]
    arch = container.assemble([f"e{i}" for i in range(len(frames))], frames, [len(b) for b in plain], hashes, [2] * len(frames))
    e = container.parse(arch).entries()
    out_size = int((e["dst_off"] + e["dst_cap"]).max())
    out, status, digest, races, ngen = run(sim, arch, e, out_size, seed=9)
    assert races == 0
    for i, b in enumerate(plain):
        assert status[i] in (0, 15, -1000), status[i]
        if status[i] == 0:
            o = int(e["dst_off"][i])
            assert np.array_equal(out[o:o + len(b)], b)
