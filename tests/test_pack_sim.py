"""The LZ4 block compressor (zpack_b200/csrc/pack_blocks.cuh) on the CPU emulation of tests/sim: every block it emits is
wrapped into an LZ4 frame and decoded by the oracle; sizes are compared with the reference algorithm's (oracle port of
LZ4F_compressFrame at level 0).  The GPU tests (tests/test_gpu_pack.py) cover the real library."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from zpack_b200 import corpus

import crafted

HERE = os.path.dirname(os.path.abspath(__file__))
SIM = os.path.join(HERE, "sim")
LIB = os.path.join(SIM, "libpack_sim.so")
CSRC = os.path.join(os.path.dirname(HERE), "zpack_b200", "csrc")


def _build():
    deps = [os.path.join(SIM, f) for f in ("pack_sim.cpp", "sim_rt.h", "sim_cuda.h")]
    deps += [os.path.join(CSRC, f) for f in ("pack_blocks.cuh", "ptx.cuh", "common.cuh")]
    if os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-DZPB_SIM", "-shared", "-fPIC", "-w",
                    os.path.join(SIM, "pack_sim.cpp"), "-o", LIB], check=True)


@pytest.fixture(scope="module")
def sim():
    _build()
    lib = C.CDLL(LIB)
    vp = C.c_void_p
    lib.sim_lz4_pack_blocks.argtypes = [vp, vp, vp, C.c_uint32, vp, vp, C.c_int, C.c_uint64, vp]
    return lib


def pack_blocks(sim, bufs, seed=1, grid=2, misalign=0, want_winop=False):
    """bufs: list of byte arrays (each <= 64 KB) -> list of (csize, payload[, window offsets])"""
    offs, total = [], misalign
    for b in bufs:
        offs.append(total)
        total += len(b) + (3 if misalign else 0)
    inp = np.zeros(total + 64, np.uint8)
    for o, b in zip(offs, bufs):
        inp[o:o + len(b)] = b
    n = len(bufs)
    so = np.array(offs, np.uint64)
    ln = np.array([len(b) for b in bufs], np.uint32)
    scratch = np.full(n * 65536 + 64, 0xEE, np.uint8)
    csize = np.full(n, 0xFFFFFFFF, np.uint32)
    winop = np.zeros(n * 17, np.uint32)
    sim.sim_lz4_pack_blocks(inp.ctypes.data, so.ctypes.data, ln.ctypes.data, n, scratch.ctypes.data, csize.ctypes.data, grid, seed,
                            winop.ctypes.data if want_winop else None)
    assert (scratch[n * 65536:] == 0xEE).all()
    res = [(int(csize[i]), scratch[i * 65536:i * 65536 + int(csize[i])].copy()) for i in range(n)]
    return [r + (winop[17 * i:17 * i + 17].copy(),) for i, r in enumerate(res)] if want_winop else res


def decode_block(oracle, payload, n):
    hdr = bytes(oracle.lz4f_encode_port(np.zeros(0, np.uint8), 0, independent=True)[:7])
    fr = hdr + len(payload).to_bytes(4, "little") + bytes(payload) + b"\0\0\0\0"
    return oracle.lz4f_decode_port(np.frombuffer(fr, np.uint8), n)


def check(sim, oracle, bufs, **kw):
    res = pack_blocks(sim, bufs, **kw)
    ours = ref = 0
    for i, (b, (c, payload)) in enumerate(zip(bufs, res)):
        assert c < len(b), (i, c, len(b))                      # compressed only if smaller, else 0 = stored
        if c:
            rc, out = decode_block(oracle, payload, len(b))
            assert rc == 0 and np.array_equal(out[:len(b)], b), (i, len(b), c)
        ours += c if c else len(b)
        ref += len(oracle.lz4f_encode_port(b, 0, independent=True)) - 15 if len(b) else 0
    return ours, ref


def test_corpus_blocks_decode_and_ratio(sim, oracle):
    bufs = [corpus.entry_bytes(i, 65536) for i in range(6)] + [corpus.entry_bytes(40 + i, s) for i, s in enumerate((40000, 4097, 4096, 300))]
    ours, ref = check(sim, oracle, bufs)
    print(f"\nblock compressor on zpk-synth-v1: {ours} bytes vs reference algorithm {ref} ({ref / ours:.3f} of its ratio)")
    assert ours < 1.06 * ref


def test_edge_sizes(sim, oracle):
    rng = np.random.default_rng(3)
    sizes = [1, 4, 5, 11, 12, 13, 14, 17, 31, 32, 33, 64, 127, 128, 129, 140, 255, 256, 1000, 4095, 4096, 4097, 4108, 4109, 8192, 65535, 65536]
    bufs = [np.frombuffer((b"abcdefgh" * 9000)[:s], np.uint8) for s in sizes]            # periodic: matches run to every limit
    bufs += [np.zeros(s, np.uint8) for s in sizes]                                     # overlapping matches at distance 1
    bufs += [rng.integers(0, 256, s, dtype=np.uint8) for s in (13, 100, 5000, 65536)]  # incompressible: stored
    bufs += [rng.integers(0, 4, s, dtype=np.uint8) for s in (100, 5000, 65536)]        # dense short matches
    check(sim, oracle, bufs, misalign=1)
    check(sim, oracle, bufs[:30], seed=9, grid=1)


def test_crafted_inputs(sim, oracle):
    """Plain texts of the crafted decoder cases (tests/crafted.py): long matches, overlaps, far offsets."""
    rng = np.random.default_rng(11)
    bufs = []
    for style in ("mixed", "overlap", "far", "long"):
        try:
            _fr, plain = crafted.frame(rng, [65536, 30000], style)
        except Exception:
            continue
        plain = np.frombuffer(plain, np.uint8)
        bufs += [plain[i:i + 65536] for i in range(0, len(plain), 65536)]
    assert bufs
    check(sim, oracle, bufs)


def zstd_frame(bodies, blocks):
    """A zstd frame laid out as pack_kernel.cuh does: magic, FHD 0xC0 (8-byte content size), window descriptor 0x38
    (128 KB), 3-byte block headers.  bodies[k]: the list of Compressed_Block bodies of block k, or None = Raw_Block."""
    total = sum(len(b) for b in blocks)
    fr = bytearray([0x28, 0xB5, 0x2F, 0xFD, 0xC0, 0x38]) + total.to_bytes(8, "little")
    for k, (subs, raw) in enumerate(zip(bodies, blocks)):
        last_block = k + 1 == len(blocks)
        if subs is None:
            fr += ((1 if last_block else 0) | (len(raw) << 3)).to_bytes(3, "little") + bytes(raw)
        else:
            for j, body in enumerate(subs):
                last = 1 if last_block and j + 1 == len(subs) else 0
                fr += (last | (2 << 1) | (len(body) << 3)).to_bytes(3, "little") + bytes(body)
    return np.frombuffer(bytes(fr), np.uint8)


@pytest.mark.parametrize("huffman", [0, 1, 3])     # bit 0: Huffman literals, bit 1: FSE tables fitted to the block
def test_zstd_blocks_from_the_lz4_matches_decode_with_the_oracle_and_the_reference(sim, oracle, huffman):
    """zstd_encode.cuh: every 4 KB window of the block compressor's LZ4 payload -> one zstd Compressed_Block (Huffman or
    raw literals + predefined-mode FSE sequences).  Frames must decode in the oracle's zstd port and in the UNMODIFIED
    reference's ZSTD_decompress."""
    sim.sim_zstd_encode_block.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    sim.sim_zstd_encode_block.restype = C.c_int
    rng = np.random.default_rng(5)
    blocks = [corpus.entry_bytes(i, 65536) for i in range(8)]
    blocks += [corpus.entry_bytes(20 + i, s) for i, s in enumerate((70, 300, 4097, 40000))]
    far = rng.integers(0, 256, 65536, dtype=np.uint8)            # short matches at far offsets: zstd bodies larger than the LZ4 bytes
    for k in range(300):
        p, q = int(rng.integers(40000, 65000)), int(rng.integers(0, 20000))
        far[p:p + 6] = far[q:q + 6]
    skew = (rng.geometric(0.35, 65536) % 100).astype(np.uint8)    # very skewed literals: codes that need the 11-bit limit
    skew[::97] = 120
    for q in range(4096, 65536, 256):
        skew[q:q + 96] = skew[q - 4096:q - 4000]                   # and enough matches for the block compressor to keep going
    blocks += [np.zeros(65536, np.uint8), np.frombuffer((b"abcdefgh" * 9000)[:65536], np.uint8),
               rng.integers(0, 4, 65536, dtype=np.uint8), rng.integers(0, 256, 5000, dtype=np.uint8), far, skew,
               np.frombuffer(b"x" * 20000 + bytes(rng.integers(97, 123, 30000, dtype=np.uint8)), np.uint8)]
    packed = pack_blocks(sim, blocks, want_winop=True)
    bodies, lz_total, z_total, used = [], 0, 0, set()
    for b, (c, payload, winop) in zip(blocks, packed):
        subs = None
        if c:
            slot = np.zeros(65536 + 16384 + 512 + 64, np.uint8)
            zbody = np.zeros(16, np.uint32)
            modes = np.zeros(16, np.uint32)
            pay = np.zeros(65536 + 16, np.uint8)
            pay[:c] = payload
            if sim.sim_zstd_encode_block(pay.ctypes.data, c, len(b), winop.ctypes.data, slot.ctypes.data, zbody.ctypes.data, huffman,
                                         modes.ctypes.data):
                nwin = (len(b) - 12) // 4096 + 1
                subs = []
                for w in range(nwin):
                    if zbody[w]:
                        o = int(winop[w]) + (int(winop[w]) >> 2) + 24 * w
                        subs.append(slot[o:o + int(zbody[w])].copy())
                        used.add(int(modes[w]))
        bodies.append(subs)
        lz_total += c if c else len(b)
        z_total += sum(len(x) + 3 for x in subs) if subs is not None else len(b)
    assert sum(x is not None for x in bodies) >= len(blocks) - 6          # the random blocks stay raw
    assert used == ({0, 2, 3} if huffman & 1 else {0})
    for k in range(len(blocks)):                       # one frame per block, and all of them in one frame
        fr = zstd_frame([bodies[k]], [blocks[k]])
        rc, got = oracle.zstd_decode_port(fr, len(blocks[k]))
        assert rc == 0 and np.array_equal(got[:len(blocks[k])], blocks[k]), k
    fr = zstd_frame(bodies, blocks)
    plain = np.concatenate(blocks)
    rc, got = oracle.zstd_decode_port(fr, len(plain))
    assert rc == 0 and np.array_equal(got[:len(plain)], plain)
    if oracle.have_ref():
        assert np.array_equal(oracle.zstd_decompress_ref(fr, len(plain)), plain)
        ref = sum(len(oracle.zstd_compress_ref(b, 3)) for b in blocks)
        print(f"\nzstd blocks (huffman={huffman}): {z_total} bytes (LZ4 payloads {lz_total}) vs ZSTD_compress level 3: {ref}")


def test_zstd_writer_on_random_shapes(sim, oracle):
    """Blocks of varied statistics (alphabet size, repetition distance and length, sizes that are not window multiples):
    every normalisation / table-description / Huffman-limit path the writer takes must give frames the oracle's port decodes."""
    sim.sim_zstd_encode_block.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    sim.sim_zstd_encode_block.restype = C.c_int
    rng = np.random.default_rng(77)
    blocks = []
    for t in range(12):
        n = int(rng.choice([65536, 65536, 50000, 20000, 9000, 4100]))
        alpha = int(rng.choice([2, 5, 16, 64, 120, 200]))
        b = rng.integers(0, alpha, n, dtype=np.uint8)
        if t % 3 == 1:                                   # skewed symbols
            b = (rng.zipf(1.3, n) % alpha).astype(np.uint8)
        reps = int(rng.choice([0, 50, 400, 3000]))
        for _ in range(reps):                            # repeats of random length at random (also very short) distances
            ln = int(rng.choice([4, 5, 8, 20, 70, 300]))
            p = int(rng.integers(ln + 1, max(ln + 2, n - ln)))
            d = int(rng.choice([1, 2, 3, 7, 64, 1000, 30000]))
            d = min(d, p)
            for k in range(ln):
                b[p + k] = b[p + k - d]
        blocks.append(b)
    packed = pack_blocks(sim, blocks, want_winop=True)
    encoded = 0
    for b, (c, payload, winop) in zip(blocks, packed):
        subs = None
        if c:
            slot = np.zeros(65536 + 16384 + 512 + 64, np.uint8)
            zbody, pay = np.zeros(16, np.uint32), np.zeros(65536 + 16, np.uint8)
            pay[:c] = payload
            if sim.sim_zstd_encode_block(pay.ctypes.data, c, len(b), winop.ctypes.data, slot.ctypes.data, zbody.ctypes.data, 3, None):
                nwin = (len(b) - 12) // 4096 + 1
                subs = [slot[int(winop[w]) + (int(winop[w]) >> 2) + 24 * w:][:int(zbody[w])].copy() for w in range(nwin) if zbody[w]]
                encoded += 1
        fr = zstd_frame([subs], [b])
        rc, got = oracle.zstd_decode_port(fr, len(b))
        assert rc == 0 and np.array_equal(got[:len(b)], b), (len(b), c)
        if subs is not None and oracle.have_ref():
            assert np.array_equal(oracle.zstd_decompress_ref(fr, len(b)), b)
    assert encoded >= 6
