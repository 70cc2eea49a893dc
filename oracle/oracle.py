"""ctypes access to the two checkers (TEST INFRASTRUCTURE ONLY — see oracle/README.md).

  port  liboracle.so          our plain-C restatement (oracle/*_oracle.c)
  ref   _ref/libzpack_ref.so  the unmodified reference compiled from /root/reference by oracle/Makefile

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_PATH = os.path.join(HERE, "liboracle.so")
REF_PATH = os.path.join(HERE, "_ref", "libzpack_ref.so")
REF_CLI = os.path.join(HERE, "_ref", "zpack_ref")

_port = None
_ref = None


def build(verbose: bool = False):
    """make liboracle.so (always) and _ref (only where /root/reference exists)."""
    subprocess.run(["make", "-C", HERE, "-j8", "all"], check=True,
                   stdout=None if verbose else subprocess.DEVNULL)


def _u8(a):
    return a.ctypes.data_as(C.c_void_p)


def port() -> C.CDLL:
    global _port
    if _port is None:
        if not os.path.exists(PORT_PATH):
            build()
        lib = C.CDLL(PORT_PATH)
        vp, sz = C.c_void_p, C.c_size_t
        lib.orc_xxh3_64.restype = C.c_uint64
        lib.orc_xxh3_64.argtypes = [vp, sz]
        lib.orc_xxh32.restype = C.c_uint32
        lib.orc_xxh32.argtypes = [vp, sz, C.c_uint32]
        lib.orc_lz4f_decode.argtypes = [vp, sz, vp, sz, C.POINTER(sz)]
        lib.orc_lz4f_encode.restype = sz
        lib.orc_lz4f_encode.argtypes = [vp, sz, vp, sz, C.c_int, C.c_int]
        lib.orc_lz4f_bound.restype = sz
        lib.orc_lz4f_bound.argtypes = [sz]
        lib.orc_zstd_decode.argtypes = [vp, sz, vp, sz, C.POINTER(sz)]
        lib.orc_read_entry.argtypes = [C.c_int, vp, sz, vp, sz, sz, C.c_uint64, C.POINTER(C.c_uint64)]
        _port = lib
    return _port


def have_ref() -> bool:
    return os.path.exists(REF_PATH)


def ref() -> C.CDLL:
    """The unmodified reference library (zpack_* + LZ4F_* + ZSTD_* + XXH3_*)."""
    global _ref
    if _ref is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref not built (needs /root/reference): run `make -C oracle ref`")
        lib = C.CDLL(REF_PATH)
        vp, sz = C.c_void_p, C.c_size_t
        lib.XXH3_64bits.restype = C.c_uint64
        lib.XXH3_64bits.argtypes = [vp, sz]
        lib.LZ4F_compressFrameBound.restype = sz
        lib.LZ4F_compressFrameBound.argtypes = [sz, vp]
        lib.LZ4F_compressFrame.restype = sz
        lib.LZ4F_compressFrame.argtypes = [vp, sz, vp, sz, vp]
        lib.LZ4F_compressBound.restype = sz
        lib.LZ4F_compressBound.argtypes = [sz, vp]
        lib.LZ4F_isError.argtypes = [sz]
        lib.ZSTD_compressBound.restype = sz
        lib.ZSTD_compressBound.argtypes = [sz]
        lib.ZSTD_compress.restype = sz
        lib.ZSTD_compress.argtypes = [vp, sz, vp, sz, C.c_int]
        lib.ZSTD_decompress.restype = sz
        lib.ZSTD_decompress.argtypes = [vp, sz, vp, sz]
        lib.ZSTD_isError.argtypes = [sz]
        lib.zpack_read_file.argtypes = [vp, vp, vp, sz, vp]
        lib.zpack_init_reader_memory_shared.argtypes = [vp, vp, sz]
        lib.zpack_close_reader.argtypes = [vp]
        lib.zpack_create_dctx.restype = vp
        lib.zpack_create_dctx.argtypes = [C.c_int]
        lib.zpack_free_dctx.argtypes = [C.c_int, vp]
        lib.zpack_init_writer_heap.argtypes = [vp, sz]
        lib.zpack_write_archive.argtypes = [vp, vp, C.c_uint64]
        lib.zpack_close_writer.argtypes = [vp]
        _ref = lib
    return _ref


# ---- struct layouts of /root/reference/lib/zpack.h (ABI) ------------------------------------
class FileEntry(C.Structure):  # zpack.h:71-80
    _fields_ = [("filename", C.c_char_p), ("offset", C.c_uint64), ("comp_size", C.c_uint64),
                ("uncomp_size", C.c_uint64), ("hash", C.c_uint64), ("comp_method", C.c_uint8)]


class Reader(C.Structure):  # zpack.h:85-110
    _fields_ = [("version", C.c_uint16), ("file_entries", C.POINTER(FileEntry)), ("file_count", C.c_uint64),
                ("comp_size", C.c_uint64), ("uncomp_size", C.c_uint64), ("file_size", C.c_size_t),
                ("zstd_dctx", C.c_void_p), ("lz4f_dctx", C.c_void_p), ("last_return", C.c_size_t),
                ("cdr_offset", C.c_uint64), ("eocdr_offset", C.c_uint64), ("buffer", C.c_void_p),
                ("buffer_shared", C.c_uint8), ("file", C.c_void_p)]


class CompressOptions(C.Structure):  # zpack.h:115-120
    _fields_ = [("method", C.c_int), ("level", C.c_int)]


class ZFile(C.Structure):  # zpack.h:125-134
    _fields_ = [("filename", C.c_char_p), ("buffer", C.c_void_p), ("size", C.c_uint64),
                ("options", C.POINTER(CompressOptions)), ("cctx", C.c_void_p)]


class Writer(C.Structure):  # zpack.h:139-164
    _fields_ = [("buffer", C.c_void_p), ("buffer_capacity", C.c_size_t), ("file", C.c_void_p),
                ("file_size", C.c_size_t), ("write_offset", C.c_size_t), ("file_entries", C.POINTER(FileEntry)),
                ("fe_capacity", C.c_uint64), ("file_count", C.c_uint64), ("zstd_cctx", C.c_void_p),
                ("lz4f_cctx", C.c_void_p), ("last_return", C.c_size_t), ("cdr_offset", C.c_uint64),
                ("eocdr_offset", C.c_uint64)]


class Stream(C.Structure):  # zpack.h:169-184
    _fields_ = [("next_in", C.c_void_p), ("avail_in", C.c_size_t), ("total_in", C.c_size_t),
                ("next_out", C.c_void_p), ("avail_out", C.c_size_t), ("total_out", C.c_size_t),
                ("read_back", C.c_size_t), ("xxh3_state", C.c_void_p)]


# ---- convenience wrappers ---------------------------------------------------------------------
def xxh3_port(data) -> int:
    a = np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data)
    return int(port().orc_xxh3_64(_u8(a) if len(a) else None, len(a)))


def xxh3_ref(data) -> int:
    a = np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data)
    return int(ref().XXH3_64bits(_u8(a) if len(a) else None, len(a)))


def lz4f_decode_port(comp, cap: int):
    a = np.ascontiguousarray(np.frombuffer(comp, np.uint8) if not isinstance(comp, np.ndarray) else comp)
    out = np.zeros(max(cap, 1), np.uint8)
    n = C.c_size_t(0)
    rc = port().orc_lz4f_decode(_u8(a), len(a), _u8(out), cap, C.byref(n))
    return rc, out[:n.value]


def lz4f_encode_port(data, level: int = 0, independent: bool = False) -> np.ndarray:
    a = np.ascontiguousarray(np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else data)
    cap = len(a) + 4 * ((len(a) + 65535) // 65536) + 32
    out = np.empty(cap, np.uint8)
    n = port().orc_lz4f_encode(_u8(a) if len(a) else None, len(a), _u8(out), cap, level, int(independent))
    assert n > 0
    return out[:n].copy()


def zstd_decode_port(comp, cap: int):
    a = np.ascontiguousarray(np.frombuffer(comp, np.uint8) if not isinstance(comp, np.ndarray) else comp)
    out = np.zeros(max(cap, 1), np.uint8)
    n = C.c_size_t(0)
    rc = port().orc_zstd_decode(_u8(a), len(a), _u8(out), cap, C.byref(n))
    return rc, out[:n.value]


def read_entry_port(method: int, comp, max_size: int, uncomp_size: int, expect_hash: int):
    a = np.ascontiguousarray(np.frombuffer(comp, np.uint8) if not isinstance(comp, np.ndarray) else comp)
    out = np.zeros(max(max_size, 1), np.uint8)
    dg = C.c_uint64(0)
    rc = port().orc_read_entry(method, _u8(a) if len(a) else None, len(a), _u8(out), max_size, uncomp_size,
                               expect_hash, C.byref(dg))
    return rc, out[:uncomp_size], dg.value


class LZ4FPrefs(C.Structure):
    """LZ4F_preferences_t (externals/lz4/lib/lz4frame.h:172-190)."""
    _fields_ = [("blockSizeID", C.c_int), ("blockMode", C.c_int), ("contentChecksumFlag", C.c_int),
                ("frameType", C.c_int), ("contentSize", C.c_ulonglong), ("dictID", C.c_uint),
                ("blockChecksumFlag", C.c_int), ("compressionLevel", C.c_int), ("autoFlush", C.c_uint),
                ("favorDecSpeed", C.c_uint), ("reserved", C.c_uint * 3)]


def lz4f_compress_ref(data, level: int = 0, block_mode: int = 0, block_size_id: int = 0,
                      content_checksum: int = 0, block_checksum: int = 0, content_size: int = 0) -> np.ndarray:
    """LZ4F_compressFrame of the unmodified reference; zeroed prefs + level = what ZPack writes."""
    a = np.ascontiguousarray(np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else data)
    p = LZ4FPrefs()
    p.compressionLevel, p.blockMode, p.blockSizeID = level, block_mode, block_size_id
    p.contentChecksumFlag, p.blockChecksumFlag, p.contentSize = content_checksum, block_checksum, content_size
    lib = ref()
    cap = lib.LZ4F_compressFrameBound(len(a), C.byref(p))
    out = np.empty(cap, np.uint8)
    n = lib.LZ4F_compressFrame(_u8(out), cap, _u8(a) if len(a) else None, len(a), C.byref(p))
    assert not lib.LZ4F_isError(n)
    return out[:n].copy()


def zstd_compress_ref(data, level: int = 3) -> np.ndarray:
    a = np.ascontiguousarray(np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else data)
    lib = ref()
    cap = lib.ZSTD_compressBound(len(a))
    out = np.empty(cap, np.uint8)
    n = lib.ZSTD_compress(_u8(out), cap, _u8(a) if len(a) else None, len(a), level)
    assert not lib.ZSTD_isError(n)
    return out[:n].copy()


def zstd_decompress_ref(comp, size: int) -> np.ndarray:
    """ZSTD_decompress of the unmodified reference (zstd 1.5.0)."""
    a = np.ascontiguousarray(np.frombuffer(comp, np.uint8) if not isinstance(comp, np.ndarray) else comp)
    lib = ref()
    lib.ZSTD_decompress.restype = C.c_size_t
    lib.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    out = np.zeros(max(size, 1), np.uint8)
    n = lib.ZSTD_decompress(_u8(out), size, _u8(a), len(a))
    assert not lib.ZSTD_isError(n), "reference zstd decoder rejects the frame"
    return out[:n].copy()


def write_archive_ref(names, buffers, method: int, level: int) -> np.ndarray:
    """zpack_write_archive into a heap writer (lib/zpack_write.c:818) — the reference's own packer."""
    lib = ref()
    n = len(names)
    opts = CompressOptions(method, level)
    files = (ZFile * n)()
    keep = []
    for i, (nm, b) in enumerate(zip(names, buffers)):
        a = np.ascontiguousarray(b)
        keep.append(a)
        files[i].filename = nm.encode()
        files[i].buffer = a.ctypes.data
        files[i].size = len(a)
        files[i].options = C.pointer(opts)
        files[i].cctx = None
    w = Writer()
    rc = lib.zpack_init_writer_heap(C.byref(w), 0)
    assert rc == 0
    rc = lib.zpack_write_archive(C.byref(w), files, n)
    if rc != 0:
        lib.zpack_close_writer(C.byref(w))
        raise RuntimeError(f"zpack_write_archive failed: {rc}")
    out = np.ctypeslib.as_array(C.cast(w.buffer, C.POINTER(C.c_uint8)), shape=(w.file_size,)).copy()
    lib.zpack_close_writer(C.byref(w))
    return out


class RefReader:
    """zpack_reader over a shared memory buffer (zpack_init_reader_memory_shared, lib/zpack.h:455)."""

    def __init__(self, archive: np.ndarray):
        self.lib = ref()
        self.buf = np.ascontiguousarray(archive)
        self.r = Reader()
        rc = self.lib.zpack_init_reader_memory_shared(C.byref(self.r), self.buf.ctypes.data, len(self.buf))
        if rc != 0:
            raise RuntimeError(f"zpack_init_reader_memory_shared failed: {rc}")
        self.count = int(self.r.file_count)

    @staticmethod
    def open_result(archive: np.ndarray) -> int:
        """zpack_result of opening these bytes with the unmodified reference (lib/zpack_read.c:225-260)"""
        lib, buf, r = ref(), np.ascontiguousarray(archive), Reader()
        rc = lib.zpack_init_reader_memory_shared(C.byref(r), buf.ctypes.data, len(buf))
        lib.zpack_close_reader(C.byref(r))
        return int(rc)

    def entry(self, i: int) -> FileEntry:
        return self.r.file_entries[i]

    def read(self, i: int, max_size=None, dctx=None):
        e = self.r.file_entries[i]
        cap = int(e.uncomp_size) if max_size is None else max_size
        out = np.zeros(max(cap, 1), np.uint8)
        rc = self.lib.zpack_read_file(C.byref(self.r), C.byref(e), out.ctypes.data, cap, dctx)
        return rc, out[:cap]

    def close(self):
        if self.r is not None:
            self.lib.zpack_close_reader(C.byref(self.r))
            self.r = None
