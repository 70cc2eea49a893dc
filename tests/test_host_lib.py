"""The lib/zpack.h drop-in (zpack_b200/libzpack.so).

CPU part (no GPU): every one of the reference's 52 entry points is exported, struct layouts equal the
reference header's, and the container-only paths behave like tests/open_archive.c expects.
GPU part: the reference's OWN test programs and CLI — compiled unmodified against the reference header by
oracle/Makefile and linked to our library (oracle/_ref/dropin_*) — run against the GPU path."""
import ctypes as C
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "zpack_b200", "libzpack.so")
REFDIR = os.path.join(ROOT, "oracle", "_ref")
REF_HEADER = "/root/reference/lib/zpack.h"
GOLD_HASHES = [0x7874CBA47D02B07D, 0x15F25C0F24DD8E52]


def _lib():
    if not os.path.exists(LIB):
        import __graft_entry__
        __graft_entry__.build()
    return C.CDLL(LIB)


def _api_names():
    text = open(os.path.join(ROOT, "zpack_b200", "host", "zpack_api.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zpack_[a-z0-9_]+)\s*\(", text)))


def test_exports_every_entry_point():
    lib = _lib()
    names = _api_names()
    assert len(names) == 53                      # the reference's 52 + zpack_read_files (batched read)
    for n in names:
        assert hasattr(lib, n), n
    if os.path.exists(REF_HEADER):
        ref = re.sub(r"/\*.*?\*/|//[^\n]*", "", open(REF_HEADER).read(), flags=re.S)
        ref_names = set(re.findall(r"ZPACK_EXPORT\s+[^;(]*?\b(zpack_[a-z0-9_]+)\s*\(", ref)) - {"zpack_convert_wchar_to_utf8"}
        assert ref_names <= set(names), ref_names - set(names)
        assert len(ref_names) == 52


LAYOUT_PROBE = r"""
#include <stdio.h>
#include <stddef.h>
#include "%s"
#define F(t, f) printf(#t "." #f " %%zu\n", offsetof(t, f));
int main(void) {
  printf("sizes %%zu %%zu %%zu %%zu %%zu %%zu\n", sizeof(zpack_file_entry), sizeof(zpack_reader), sizeof(zpack_compress_options),
         sizeof(zpack_file), sizeof(zpack_writer), sizeof(zpack_stream));
  F(zpack_file_entry, offset) F(zpack_file_entry, hash) F(zpack_file_entry, comp_method)
  F(zpack_reader, file_entries) F(zpack_reader, file_size) F(zpack_reader, lz4f_dctx) F(zpack_reader, last_return)
  F(zpack_reader, cdr_offset) F(zpack_reader, buffer) F(zpack_reader, buffer_shared) F(zpack_reader, file)
  F(zpack_file, buffer) F(zpack_file, size) F(zpack_file, options) F(zpack_file, cctx)
  F(zpack_writer, buffer_capacity) F(zpack_writer, file) F(zpack_writer, write_offset) F(zpack_writer, file_entries)
  F(zpack_writer, file_count) F(zpack_writer, last_return) F(zpack_writer, eocdr_offset)
  F(zpack_stream, avail_in) F(zpack_stream, next_out) F(zpack_stream, total_out) F(zpack_stream, read_back) F(zpack_stream, xxh3_state)
  printf("enum %%d %%d %%d %%d\n", ZPACK_ERROR_BUFFER_TOO_SMALL, ZPACK_ERROR_FILE_HASH_MISMATCH, ZPACK_ERROR_STREAM_INVALID, ZPACK_ERROR_NOT_AVAILABLE);
  return 0; }
"""


def test_struct_layouts_equal_the_reference_header(tmp_path):
    if not os.path.exists(REF_HEADER):
        pytest.skip("/root/reference absent")
    outs = []
    for tag, hdr in (("ref", REF_HEADER), ("ours", os.path.join(ROOT, "zpack_b200", "host", "zpack_api.h"))):
        src = tmp_path / f"probe_{tag}.c"
        src.write_text(LAYOUT_PROBE % hdr)
        exe = tmp_path / f"probe_{tag}"
        subprocess.run(["gcc", "-w", "-o", str(exe), str(src)], check=True)
        outs.append(subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout)
    assert outs[0] == outs[1]
    assert "enum 12 15 21 24" in outs[1]


class FileEntry(C.Structure):
    _fields_ = [("filename", C.c_char_p), ("offset", C.c_uint64), ("comp_size", C.c_uint64),
                ("uncomp_size", C.c_uint64), ("hash", C.c_uint64), ("comp_method", C.c_uint8)]


class Reader(C.Structure):
    _fields_ = [("version", C.c_uint16), ("file_entries", C.POINTER(FileEntry)), ("file_count", C.c_uint64),
                ("comp_size", C.c_uint64), ("uncomp_size", C.c_uint64), ("file_size", C.c_size_t),
                ("zstd_dctx", C.c_void_p), ("lz4f_dctx", C.c_void_p), ("last_return", C.c_size_t),
                ("cdr_offset", C.c_uint64), ("eocdr_offset", C.c_uint64), ("buffer", C.c_void_p),
                ("buffer_shared", C.c_uint8), ("file", C.c_void_p)]


class Options(C.Structure):
    _fields_ = [("method", C.c_int), ("level", C.c_int)]


class ZFile(C.Structure):
    _fields_ = [("filename", C.c_char_p), ("buffer", C.c_void_p), ("size", C.c_uint64),
                ("options", C.POINTER(Options)), ("cctx", C.c_void_p)]


class Writer(C.Structure):
    _fields_ = [("buffer", C.c_void_p), ("buffer_capacity", C.c_size_t), ("file", C.c_void_p),
                ("file_size", C.c_size_t), ("write_offset", C.c_size_t), ("file_entries", C.POINTER(FileEntry)),
                ("fe_capacity", C.c_uint64), ("file_count", C.c_uint64), ("zstd_cctx", C.c_void_p),
                ("lz4f_cctx", C.c_void_p), ("last_return", C.c_size_t), ("cdr_offset", C.c_uint64),
                ("eocdr_offset", C.c_uint64)]


class Stream(C.Structure):
    _fields_ = [("next_in", C.c_void_p), ("avail_in", C.c_size_t), ("total_in", C.c_size_t),
                ("next_out", C.c_void_p), ("avail_out", C.c_size_t), ("total_out", C.c_size_t),
                ("read_back", C.c_size_t), ("xxh3_state", C.c_void_p)]


@pytest.mark.parametrize("kind", ["none", "zstd", "lz4"])
def test_open_archive_three_ways_like_the_reference_test(golden_dir, kind):
    """tests/open_archive.c:21-24,59-85: path, copied buffer, shared buffer; names, sizes, digests."""
    lib = _lib()
    path = os.path.join(golden_dir, f"archive_{kind}.zpk")
    raw = np.fromfile(path, np.uint8)
    for how in ("path", "copy", "shared"):
        r = Reader()
        if how == "path":
            rc = lib.zpack_init_reader(C.byref(r), path.encode())
        elif how == "copy":
            rc = lib.zpack_init_reader_memory(C.byref(r), raw.ctypes.data_as(C.c_void_p), C.c_size_t(len(raw)))
        else:
            rc = lib.zpack_init_reader_memory_shared(C.byref(r), raw.ctypes.data_as(C.c_void_p), C.c_size_t(len(raw)))
        assert rc == 0
        assert r.file_count == 2 and r.version == 1
        assert [r.file_entries[i].filename for i in range(2)] == [b"file1.txt", b"file2.txt"]
        assert [r.file_entries[i].uncomp_size for i in range(2)] == [169, 349]
        assert [r.file_entries[i].hash for i in range(2)] == GOLD_HASHES
        assert r.file_entries[0].comp_method == {"none": 0, "zstd": 1, "lz4": 2}[kind]
        lib.zpack_get_file_entry.restype = C.POINTER(FileEntry)
        lib.zpack_get_file_entry.argtypes = [C.c_char_p, C.POINTER(FileEntry), C.c_uint64]
        assert lib.zpack_get_file_entry(b"file2.txt", r.file_entries, r.file_count).contents.uncomp_size == 349
        assert not lib.zpack_get_file_entry(b"nope", r.file_entries, r.file_count)
        lib.zpack_close_reader(C.byref(r))
        assert r.file_count == 0 and not r.buffer


def test_container_errors_and_stream_sizes():
    lib = _lib()
    r = Reader()
    junk = np.zeros(100, np.uint8)
    assert lib.zpack_init_reader_memory_shared(C.byref(r), junk.ctypes.data_as(C.c_void_p), C.c_size_t(100)) == 6   # SIGNATURE_INVALID
    r = Reader()
    assert lib.zpack_init_reader_memory_shared(C.byref(r), junk.ctypes.data_as(C.c_void_p), C.c_size_t(10)) == 5    # FILE_TOO_SMALL
    assert lib.zpack_init_reader(C.byref(Reader()), b"/nonexistent/x.zpk") == 3                                      # OPEN_FAILED
    for fn, want in (("zpack_get_dstream_in_size", [131075, 131075, 65551, 0]), ("zpack_get_dstream_out_size", [131072, 131072, 65536, 0]),
                     ("zpack_get_cstream_in_size", [131072, 131072, 65536, 0]), ("zpack_get_cstream_out_size", [131591, 131591, 65551, 0])):
        f = getattr(lib, fn)
        f.restype = C.c_size_t
        assert [f(m) for m in range(4)] == want          # the values the reference returns (lib/zpack_read.c:719-758)


# ------------------------------------------------------------------------------------------------ GPU
def _workdir(tmp_path, golden_dir):
    for f in os.listdir(golden_dir):
        if f.endswith((".zpk", ".txt")):
            shutil.copy(os.path.join(golden_dir, f), tmp_path / f)
    return tmp_path


def _have(exe):
    return os.path.exists(os.path.join(REFDIR, exe))


@pytest.mark.gpu
def test_reference_read_archive_program_on_the_dropin(tmp_path, golden_dir):
    """tests/read_archive.c, unmodified: one-shot + 16-byte-buffer streaming reads, from file and from memory."""
    if not _have("dropin_read_archive"):
        pytest.skip("oracle/_ref/dropin_read_archive not built (needs /root/reference at build time)")
    wd = _workdir(tmp_path, golden_dir)
    r = subprocess.run([os.path.join(REFDIR, "dropin_read_archive")], cwd=wd, capture_output=True, text=True, timeout=300)
    sections = r.stdout.split("Archive #")
    by_name = {s.split("(")[1].split(")")[0]: s for s in sections[1:]}
    for name in ("archive_none.zpk", "archive_lz4.zpk"):
        s = by_name[name]
        assert "Failed" not in s and "invalid" not in s, s
        assert s.count("is valid") == 8, s            # 2 files x (one-shot + streaming) x (file + buffer)
    z = by_name["archive_zstd.zpk"]
    assert "Failed" not in z and "invalid" not in z and "error" not in z.lower(), z
    assert z.count("is valid") == 8, z


@pytest.mark.gpu
def test_reference_open_archive_program_on_the_dropin(tmp_path, golden_dir):
    """tests/open_archive.c, unmodified, linked against libzpack.so: the three golden archives open from file and memory."""
    if not _have("dropin_open_archive"):
        pytest.skip("oracle/_ref/dropin_open_archive not built (needs /root/reference at build time)")
    wd = _workdir(tmp_path, golden_dir)
    r = subprocess.run([os.path.join(REFDIR, "dropin_open_archive")], cwd=wd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    assert "Failed" not in r.stdout and "rror" not in r.stdout, r.stdout[-2000:]


@pytest.mark.gpu
def test_reference_write_archive_program_on_the_dropin(tmp_path, golden_dir, oracle):
    """tests/write_archive.c, unmodified (/root/reference/tests/write_archive.c:100-186): every method — none, zstd, lz4 —
    written four ways (file / heap x one-shot / streaming).  The program checks return codes only, so the archives it left
    behind are then opened by the UNMODIFIED reference reader (`zpack_ref t`), which must find no corrupted file."""
    if not (_have("dropin_write_archive") and _have("zpack_ref")):
        pytest.skip("oracle/_ref/dropin_write_archive not built (needs /root/reference at build time)")
    wd = _workdir(tmp_path, golden_dir)
    r = subprocess.run([os.path.join(REFDIR, "dropin_write_archive")], cwd=wd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-2000:])
    assert r.stdout.count("Archive write successful") == 12, r.stdout[-3000:]     # 3 methods x 4 ways
    written = sorted(f for f in os.listdir(wd) if f.startswith("out_") and f.endswith(".zpk"))
    assert len(written) == 6, written                                               # the file-backed ones
    for f in written:
        t = subprocess.run([os.path.join(REFDIR, "zpack_ref"), "t", f], cwd=wd, capture_output=True, text=True, timeout=300)
        assert t.returncode == 0 and "Corrupted files: 0/" in t.stdout, (f, t.stdout[-1000:])


@pytest.mark.gpu
def test_reference_cli_on_the_dropin_round_trips_with_the_reference_cli(tmp_path, oracle):
    """BASELINE config C1 in miniature: the unmodified CLI linked to our library packs (LZ4, streamed writes) and
    extracts; the unmodified reference CLI extracts our archive and vice versa; all trees identical."""
    if not (_have("dropin_zpack") and _have("zpack_ref")):
        pytest.skip("CLI binaries not built")
    from zpack_b200 import corpus
    src = tmp_path / "corpus"
    src.mkdir()
    for i in range(24):
        d = src / corpus.CLASSES[i & 3]
        d.mkdir(exist_ok=True)
        corpus.entry_bytes(i, [65536, 1000, 200000, 0][i % 4] if i % 5 else 65536).tofile(d / f"{i:04d}.bin")
    ours, ref = os.path.join(REFDIR, "dropin_zpack"), os.path.join(REFDIR, "zpack_ref")

    def run(*a):
        r = subprocess.run(list(a), cwd=tmp_path, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, (a, r.stdout[-2000:], r.stderr[-2000:])
        return r.stdout

    run(ours, "c", "-m", "lz4", "gpu.zpk", "corpus")
    run(ref, "c", "-m", "lz4", "ref.zpk", "corpus")
    assert "Corrupted files: 0/" in run(ref, "t", "gpu.zpk")          # reference reader accepts the GPU-written archive
    assert "Corrupted files: 0/" in run(ours, "t", "ref.zpk")         # GPU reader accepts the reference-written archive
    run(ref, "x", "-o", "out_ref_from_gpu", "gpu.zpk")
    run(ours, "x", "-o", "out_gpu_from_ref", "ref.zpk")
    run(ours, "x", "-o", "out_gpu_from_gpu", "gpu.zpk")
    for out in ("out_ref_from_gpu", "out_gpu_from_ref", "out_gpu_from_gpu"):
        r = subprocess.run(["diff", "-r", str(src), str(tmp_path / out / "corpus")], capture_output=True, text=True)
        if r.returncode != 0:      # the CLI may or may not keep the top directory name
            r = subprocess.run(["diff", "-r", str(src), str(tmp_path / out)], capture_output=True, text=True)
        assert r.returncode == 0, (out, r.stdout[-1000:])


@pytest.mark.gpu
@pytest.mark.parametrize("method,level", [(0, 0), (2, 1), (1, 3)])
def test_write_archive_four_ways_and_read_back(tmp_path, oracle, method, level):
    """tests/write_archive.c:145-186 (file|heap x one-shot|streaming) for none and lz4 — and, unlike the reference
    test, every archive is read back: by our reader, and by the unmodified reference reader when present."""
    lib = _lib()
    from zpack_b200 import corpus
    bufs = [corpus.entry_bytes(i, s) for i, s in enumerate([169, 349, 70000, 0, 131072])]
    names = [f"f{i}.bin".encode() for i in range(len(bufs))]
    opt = Options(method, level)
    archives = []
    for sink in ("file", "heap"):
        for mode in ("oneshot", "stream"):
            w = Writer()
            path = tmp_path / f"out_{method}_{sink}_{mode}.zpk"
            assert (lib.zpack_init_writer(C.byref(w), str(path).encode()) if sink == "file"
                    else lib.zpack_init_writer_heap(C.byref(w), C.c_size_t(0))) == 0
            if mode == "oneshot":
                files = (ZFile * len(bufs))()
                for i, b in enumerate(bufs):
                    files[i].filename, files[i].buffer, files[i].size = names[i], b.ctypes.data if len(b) else None, len(b)
                    files[i].options = C.pointer(opt)
                assert lib.zpack_write_archive(C.byref(w), files, C.c_uint64(len(bufs))) == 0
            else:
                assert lib.zpack_write_header(C.byref(w)) == 0 and lib.zpack_write_data_header(C.byref(w)) == 0
                s = Stream()
                assert lib.zpack_init_stream(C.byref(s)) == 0
                scratch = np.zeros(65551, np.uint8)
                for i, b in enumerate(bufs):
                    lib.zpack_reset_stream(C.byref(s))
                    s.next_out, s.avail_out = scratch.ctypes.data, len(scratch)
                    for at in range(0, len(b), 65536):
                        chunk = np.ascontiguousarray(b[at:at + 65536])
                        s.next_in, s.avail_in = chunk.ctypes.data, len(chunk)
                        assert lib.zpack_write_file_stream(C.byref(w), C.byref(opt), C.byref(s), None) == 0
                    assert lib.zpack_write_file_stream_end(C.byref(w), names[i], C.byref(opt), C.byref(s), None) == 0
                    assert s.total_in == len(b)
                lib.zpack_close_stream(C.byref(s))
                assert lib.zpack_write_cdr(C.byref(w)) == 0 and lib.zpack_write_eocdr(C.byref(w)) == 0
            if sink == "heap":
                arch = np.ctypeslib.as_array(C.cast(w.buffer, C.POINTER(C.c_uint8)), shape=(w.file_size,)).copy()
            lib.zpack_close_writer(C.byref(w))
            if sink == "file":
                arch = np.fromfile(path, np.uint8)
            archives.append(arch)
    for arch in archives:
        r = Reader()
        assert lib.zpack_init_reader_memory_shared(C.byref(r), arch.ctypes.data_as(C.c_void_p), C.c_size_t(len(arch))) == 0
        assert r.file_count == len(bufs)
        for i, b in enumerate(bufs):
            e = r.file_entries[i]
            assert e.uncomp_size == len(b) and e.hash == oracle.xxh3_port(b) and e.comp_method == method
            out = np.zeros(max(len(b), 1), np.uint8)
            assert lib.zpack_read_file(C.byref(r), C.byref(e), out.ctypes.data_as(C.c_void_p), C.c_size_t(len(b)), None) == 0
            assert np.array_equal(out[:len(b)], b)
        lib.zpack_close_reader(C.byref(r))
        if oracle.have_ref():
            rd = oracle.RefReader(arch)
            for i, b in enumerate(bufs):
                rc, out = rd.read(i)
                assert rc == 0 and np.array_equal(out[:len(b)], b)
            rd.close()


@pytest.mark.gpu
def test_batched_read_extension_and_stream_with_tiny_buffers(golden_dir, oracle):
    lib = _lib()
    from zpack_b200 import container, corpus
    n, size = 64, 131072
    bufs = [corpus.entry_bytes(i, size) for i in range(n)]
    frames = [oracle.lz4f_encode_port(b, 0, False) for b in bufs]
    hashes = [oracle.xxh3_port(b) for b in bufs]
    arch = container.assemble([f"{i}" for i in range(n)], frames, [size] * n, hashes, [2] * n)
    r = Reader()
    assert lib.zpack_init_reader_memory_shared(C.byref(r), arch.ctypes.data_as(C.c_void_p), C.c_size_t(len(arch))) == 0
    out = np.zeros(n * size, np.uint8)
    off = (np.arange(n, dtype=np.uint64) * size)
    cap = np.full(n, size, np.uint64)
    st = np.full(n, -1, np.int32)
    lib.zpack_read_files.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    assert lib.zpack_read_files(C.byref(r), r.file_entries, n, out.ctypes.data, len(out), off.ctypes.data, cap.ctypes.data, st.ctypes.data) == 0
    assert (st == 0).all() and np.array_equal(out.reshape(n, size), np.stack(bufs))
    # streaming read of one entry with a 16-byte input buffer and a 350-byte output buffer (tests/read_archive.c:12,47-80)
    s = Stream()
    assert lib.zpack_init_stream(C.byref(s)) == 0
    lib.zpack_reset_stream(C.byref(s))
    inb, outb = np.zeros(16, np.uint8), np.zeros(350, np.uint8)
    got = bytearray()
    e = r.file_entries[3]
    lib.zpack_read_stream_done.restype = C.c_uint8
    for _ in range(200000):
        if s.read_back:
            C.memmove(inb.ctypes.data, s.next_in - s.read_back, s.read_back)
        s.next_in, s.avail_in = inb.ctypes.data, 16
        s.next_out, s.avail_out = outb.ctypes.data, 350
        assert lib.zpack_read_file_stream(C.byref(r), C.byref(e), C.byref(s), None) == 0
        got += bytes(outb[:350 - s.avail_out])
        if lib.zpack_read_stream_done(C.byref(s), C.byref(e)):
            break
    assert s.total_in == e.comp_size and s.total_out == size and bytes(got) == bufs[3].tobytes()
    lib.zpack_close_stream(C.byref(s))
    lib.zpack_close_reader(C.byref(r))


@pytest.mark.gpu
def test_large_entry_takes_the_block_parallel_path_and_round_trips(oracle):
    """One 12 MiB file written by zpack_write_archive (GPU writer: independent 64 KB blocks) and read back by
    zpack_read_file — comp_size >= 4 MiB, so the read goes through zpb_unpack_entry_blocks_host (frame index,
    warp per block, XXH3 chain) — and by the unmodified reference reader."""
    lib = _lib()
    from zpack_b200 import corpus
    from zpack_b200 import lib as zlib
    total = 12 << 20
    data = corpus.big_entry(total, piece=1 << 20)
    opt = Options(2, 0)
    w = Writer()
    assert lib.zpack_init_writer_heap(C.byref(w), C.c_size_t(0)) == 0
    files = (ZFile * 1)()
    files[0].filename, files[0].buffer, files[0].size, files[0].options = b"big.bin", data.ctypes.data, total, C.pointer(opt)
    assert lib.zpack_write_archive(C.byref(w), files, C.c_uint64(1)) == 0
    arch = np.ctypeslib.as_array(C.cast(w.buffer, C.POINTER(C.c_uint8)), shape=(w.file_size,)).copy()
    lib.zpack_close_writer(C.byref(w))
    r = Reader()
    assert lib.zpack_init_reader_memory_shared(C.byref(r), arch.ctypes.data_as(C.c_void_p), C.c_size_t(len(arch))) == 0
    e = r.file_entries[0]
    assert e.comp_size >= (4 << 20) and e.hash == oracle.xxh3_port(data)
    frame = arch[e.offset:e.offset + e.comp_size]
    assert zlib.lz4_frame_index(frame) is not None                  # eligible: this read is block-parallel
    out = np.zeros(total, np.uint8)
    assert lib.zpack_read_file(C.byref(r), C.byref(e), out.ctypes.data_as(C.c_void_p), C.c_size_t(total), None) == 0
    assert np.array_equal(out, data)
    # a flipped payload byte inside a stored (random) block: decode succeeds, digest does not (lib/zpack_read.c:466-468)
    blocks, _, _ = zlib.lz4_frame_index(frame)
    k = int(np.flatnonzero(blocks["flags"] == zlib.BLK_STORED)[0])
    arch[e.offset + int(blocks["src_off"][k]) + 5] ^= 0x40
    assert lib.zpack_read_file(C.byref(r), C.byref(e), out.ctypes.data_as(C.c_void_p), C.c_size_t(total), None) == 15
    arch[e.offset + int(blocks["src_off"][k]) + 5] ^= 0x40
    lib.zpack_close_reader(C.byref(r))
    if oracle.have_ref():
        rd = oracle.RefReader(arch)
        rc, ref_out = rd.read(0)
        assert rc == 0 and np.array_equal(ref_out[:total], data)
        rd.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["memory", "file"])
def test_sequential_reads_are_served_by_the_read_ahead(tmp_path, oracle, mode):
    """The CLI's `t` / `x` pattern (programs/commands.c): zpack_read_file on reader.file_entries[0], [1], [2] ... in
    order.  From the second consecutive call on, the drop-in decodes the entries behind the one asked for in the same GPU
    call and serves the next calls from that cache; bytes, return codes and last_return must be what one call per entry
    gives: a corrupted entry in the middle still reports FILE_HASH_MISMATCH (15) on its own call only, stored / zstd /
    lz4 entries mix, an entry whose fields the caller changed after the fill is decoded again, out-of-order reads work."""
    lib = _lib()
    from zpack_b200 import container, corpus
    n = 40
    sizes = [0 if i == 7 else 1000 + 3001 * i for i in range(n)]
    bufs = [corpus.entry_bytes(i, s) for i, s in enumerate(sizes)]
    methods = [(0, 2, 1)[i % 3] if oracle.have_ref() else (0, 2)[i % 2] for i in range(n)]
    frames = []
    for b, m in zip(bufs, methods):
        frames.append(b if m == 0 else oracle.lz4f_encode_port(b, 0, False) if m == 2 else oracle.zstd_compress_ref(b, 3))
    hashes = [oracle.xxh3_port(b) for b in bufs]
    arch = container.assemble([f"f{i}" for i in range(n)], frames, sizes, hashes, methods)
    d = container.parse(arch)
    bad = next(i for i in range(10, n) if methods[i] == 2)   # an LZ4 entry: flip a literal byte inside its payload
    arch[int(d.offset[bad]) + int(d.comp_size[bad]) - 9] ^= 0x01
    r = Reader()
    if mode == "memory":
        assert lib.zpack_init_reader_memory_shared(C.byref(r), arch.ctypes.data_as(C.c_void_p), C.c_size_t(len(arch))) == 0
    else:
        p = tmp_path / "ra.zpk"
        arch.tofile(p)
        assert lib.zpack_init_reader(C.byref(r), str(p).encode()) == 0
    lib.zpack_read_file.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    esz = C.sizeof(FileEntry)
    base = C.cast(r.file_entries, C.c_void_p).value

    def read(i):
        out = np.full(max(sizes[i], 1), 0xCD, np.uint8)
        rc = lib.zpack_read_file(C.byref(r), base + i * esz, out.ctypes.data, C.c_size_t(len(out)), None)
        return rc, out[:sizes[i]]

    for order in (range(n), [5, 6, 7, 8, 9, 3, 4, 30, 31, 32, 33, 9, 10, 11, 12, 13], reversed(range(n))):
        for i in order:
            rc, out = read(i)
            if i == bad:
                assert rc in (15, 10), (i, rc)            # digest mismatch (or a decode error, if the flip broke a sequence)
            else:
                assert rc == 0 and np.array_equal(out, bufs[i]), (i, rc, mode)
    # the caller edits an entry the cache holds: the cached result must not be used for it
    for i in (20, 21, 22):
        assert read(i)[0] == 0
    r.file_entries[23].hash ^= 1
    assert read(23)[0] == 15
    r.file_entries[23].hash ^= 1
    rc, out = read(23)
    assert rc == 0 and np.array_equal(out, bufs[23])
    lib.zpack_close_reader(C.byref(r))
