"""CPU check of the product's zstd decoder CONTROL LOGIC: zpack_b200/csrc/zstd_decode.cuh compiled for the
host with a 1-lane warp (tests/sim/zstd_sim.cpp) against the oracle.  The kernel proper is checked on the
GPU by tests/test_gpu_zstd.py; this suite catches table / bit-reader / sequence-rule mistakes where there
is no GPU.  The simulation library is a test artefact — nothing under zpack_b200/ loads it."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SIM_SRC = os.path.join(HERE, "sim", "zstd_sim.cpp")
SIM_LIB = os.path.join(HERE, "sim", "libzstd_sim.so")
HDR = os.path.join(os.path.dirname(HERE), "zpack_b200", "csrc", "zstd_decode.cuh")


@pytest.fixture(scope="module")
def sim():
    if (not os.path.exists(SIM_LIB) or
            os.path.getmtime(SIM_LIB) < max(os.path.getmtime(SIM_SRC), os.path.getmtime(HDR))):
        subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-w", "-o", SIM_LIB, SIM_SRC], check=True)
    lib = C.CDLL(SIM_LIB)
    lib.zs_sim_decode.restype = C.c_int
    lib.zs_sim_decode.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]

    def decode(comp, cap):
        comp = np.ascontiguousarray(comp, np.uint8)
        out = np.zeros(max(cap, 1), np.uint8)
        n = C.c_uint64(0)
        rc = lib.zs_sim_decode(comp.ctypes.data, len(comp), out.ctypes.data, cap, C.byref(n))
        return rc, out[:n.value]
    return decode


def test_sim_decodes_reference_written_frames(sim, zstd_cases):
    """Frames written by the unmodified reference at levels 1..19 (tests/golden/make_golden.py)."""
    seen = 0
    for k, comp in zstd_cases.items():
        if k.endswith("__in"):
            continue
        want = zstd_cases[k.split("__")[0] + "__in"]
        rc, out = sim(comp, len(want))
        assert rc == 0 and np.array_equal(out, want), k
        seen += 1
    assert seen >= 10


def test_sim_golden_archive(sim, golden_dir):
    from zpack_b200 import container
    arch = np.fromfile(os.path.join(golden_dir, "archive_zstd.zpk"), np.uint8)
    d = container.parse(arch)
    for i in range(len(d)):
        comp = arch[int(d.offset[i]):int(d.offset[i] + d.comp_size[i])]
        rc, out = sim(comp, int(d.uncomp_size[i]))
        assert rc == 0 and len(out) == int(d.uncomp_size[i])


def test_sim_matches_oracle_on_errors_and_multiframe(sim, oracle, zstd_cases):
    comp, want = zstd_cases["text_5k__l3"], zstd_cases["text_5k__in"]
    skip = np.array([0x50, 0x2A, 0x4D, 0x18, 3, 0, 0, 0, 1, 2, 3], np.uint8)
    cases = [(comp[:-1], len(want)), (comp, len(want) - 1), (np.concatenate([comp, np.zeros(2, np.uint8)]), len(want)),
             (np.concatenate([comp, skip, zstd_cases["one__l3"]]), len(want) + 1), (np.zeros(0, np.uint8), 10)]
    for c, cap in cases:
        rc_o, out_o = oracle.zstd_decode_port(c, cap)
        rc_s, out_s = sim(c, cap)
        assert rc_o == rc_s
        if rc_o == 0:
            assert np.array_equal(out_o, out_s)


def test_sim_matches_oracle_on_corruption(sim, oracle, zstd_cases):
    """Bit flips and truncations: the simulated kernel logic and the oracle must give the same verdict, and the
    same bytes whenever both accept."""
    rng = np.random.default_rng(5)
    keys = [k for k in zstd_cases if not k.endswith("__in")]
    n_ok = n_bad = 0
    for k in keys:
        comp = zstd_cases[k]
        cap = len(zstd_cases[k.split("__")[0] + "__in"])
        for trial in range(40):
            m = comp.copy()
            if len(m) > 1 and trial % 5 == 4:
                m = m[:int(rng.integers(1, len(m)))]
            else:
                m[int(rng.integers(0, len(m)))] ^= 1 << int(rng.integers(0, 8))
            rc_o, out_o = oracle.zstd_decode_port(m, cap)
            rc_s, out_s = sim(m, cap)
            assert rc_o == rc_s, (k, trial)
            if rc_o == 0:
                assert np.array_equal(out_o, out_s), (k, trial)
                n_ok += 1
            else:
                n_bad += 1
    assert n_bad > 100


def test_sim_on_synthetic_corpus(sim, oracle):
    """The C4 entry shape (128 KiB, level 3, all four classes) + multi-block and other levels, packed by the
    unmodified reference where it is present (this container)."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not present")
    from zpack_b200 import corpus
    for i, (size, lvl) in enumerate([(131072, 3), (131072, 3), (131072, 3), (131072, 3), (300000, 1), (70000, 5),
                                     (65536, 19), (1000, 3), (200000, 9), (13, 3), (131072, 1), (524288, 3)]):
        data = corpus.entry_bytes(i, size)
        comp = oracle.zstd_compress_ref(data, lvl)
        rc, out = sim(comp, size)
        assert rc == 0 and np.array_equal(out, data), (i, size, lvl)
