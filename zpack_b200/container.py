"""ZPack container framing v1, host side (header, data signature, CDR, EOCDR).

Format: /root/reference/docs/specs.md:18-79; writer /root/reference/lib/zpack_write.c:687-711,
778-785; reader /root/reference/lib/zpack_read.c:33-166,225-260.  All integers little-endian.
This is the "compressed-size offset table assembled on the host" of the north-star: the GPU
returns comp_size[] per entry, the exclusive prefix sum that yields entry.offset happens here.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import List, Sequence

import numpy as np

from .lib import Entry

SIG_HEADER, SIG_DATA, SIG_CDR, SIG_EOCDR = 0x154B505A, 0x144B505A, 0x134B505A, 0x124B505A
HEADER_SIZE, SIG_SIZE, CDR_HEADER_SIZE, ENTRY_FIXED, EOCDR_SIZE = 6, 4, 20, 35, 12
DATA_START = HEADER_SIZE + SIG_SIZE
MIN_ARCHIVE = HEADER_SIZE + SIG_SIZE + CDR_HEADER_SIZE + EOCDR_SIZE


class ArchiveError(ValueError):
    pass


@dataclass
class Directory:
    """Parsed central directory: one row per entry (the GPU batch descriptor source)."""
    names: List[str]
    offset: np.ndarray       # u64
    comp_size: np.ndarray    # u64
    uncomp_size: np.ndarray  # u64
    hash: np.ndarray         # u64
    method: np.ndarray       # u8
    cdr_offset: int
    file_size: int

    def __len__(self):
        return len(self.names)

    def entries(self, dst_off=None, dst_cap=None, align: int = 16) -> np.ndarray:
        """zpb_entry table; outputs laid out back to back (each slot `align`-byte aligned)."""
        n = len(self)
        e = np.zeros(n, Entry)
        e["src_off"], e["comp_size"], e["uncomp_size"] = self.offset, self.comp_size, self.uncomp_size
        e["hash"], e["method"] = self.hash, self.method
        cap = self.uncomp_size if dst_cap is None else np.asarray(dst_cap, np.uint64)
        e["dst_cap"] = cap
        if dst_off is None:
            padded = (cap + np.uint64(align - 1)) & ~np.uint64(align - 1)
            dst_off = np.concatenate([[0], np.cumsum(padded)[:-1]]).astype(np.uint64) if n else np.zeros(0, np.uint64)
        e["dst_off"] = dst_off
        return e


def parse(buf) -> Directory:
    """Open an archive held in memory (zpack_read_archive_memory, lib/zpack_read.c:225-260)."""
    a = np.frombuffer(buf, np.uint8) if not isinstance(buf, np.ndarray) else buf
    size = len(a)
    if size < MIN_ARCHIVE:
        raise ArchiveError("file too small")
    mv = memoryview(a)
    sig, ver = struct.unpack_from("<IH", mv, 0)
    if sig != SIG_HEADER:
        raise ArchiveError("bad header signature")
    if ver != 1:
        raise ArchiveError("unsupported version")
    if struct.unpack_from("<I", mv, HEADER_SIZE)[0] != SIG_DATA:
        raise ArchiveError("bad data signature")
    esig, cdr_off = struct.unpack_from("<IQ", mv, size - EOCDR_SIZE)
    if esig != SIG_EOCDR:
        raise ArchiveError("bad EOCDR signature")
    if cdr_off >= size or cdr_off + CDR_HEADER_SIZE > size - EOCDR_SIZE:
        raise ArchiveError("CDR offset out of range")
    csig, count, block = struct.unpack_from("<IQQ", mv, cdr_off)
    if csig != SIG_CDR:
        raise ArchiveError("bad CDR signature")
    if CDR_HEADER_SIZE + block > size - cdr_off or count * ENTRY_FIXED > block:
        raise ArchiveError("bad CDR block size")
    names, fixed = [], np.empty((count, 33), np.uint8)
    p, left = cdr_off + CDR_HEADER_SIZE, block
    for i in range(count):
        if left < 2:
            raise ArchiveError("bad CDR block size")
        (nl,) = struct.unpack_from("<H", mv, p)
        if ENTRY_FIXED + nl > left:
            raise ArchiveError("bad CDR block size")
        names.append(bytes(mv[p + 2:p + 2 + nl]).decode("utf-8", "surrogateescape"))
        fixed[i] = a[p + 2 + nl:p + 2 + nl + 33]
        p += ENTRY_FIXED + nl
        left -= ENTRY_FIXED + nl
    q = np.ascontiguousarray(fixed[:, :32]).view("<u8").reshape(count, 4) if count else np.zeros((0, 4), "<u8")
    return Directory(names, q[:, 0].copy(), q[:, 1].copy(), q[:, 2].copy(), q[:, 3].copy(),
                     fixed[:, 32].copy(), int(cdr_off), size)


def directory_from_table(table: np.ndarray, names_blob, file_size: int) -> Directory:
    """Directory from what zpb_archive_open_device returns (lib.Context.archive_open_device): the ArcEntry table parsed on
    the device and the directory block its name_off / name_len point into — the same object parse() builds on the host."""
    blob = bytes(memoryview(np.ascontiguousarray(names_blob, np.uint8)))
    names = [blob[int(o):int(o) + int(k)].decode("utf-8", "surrogateescape") for o, k in zip(table["name_off"], table["name_len"])]
    block = len(blob)
    return Directory(names, table["offset"].astype(np.uint64), table["comp_size"].astype(np.uint64),
                     table["uncomp_size"].astype(np.uint64), table["hash"].astype(np.uint64), table["method"].astype(np.uint8),
                     file_size - EOCDR_SIZE - CDR_HEADER_SIZE - block, file_size)


def cdr_bytes(names: Sequence[str], offset, comp, uncomp, hashes, method) -> bytes:
    """Central directory record for the given entries (zpack_write_cdr_memory, zpack_write.c:687-711)."""
    n = len(names)
    enc = [s.encode("utf-8", "surrogateescape") for s in names]
    for b in enc:
        if len(b) > 65535:
            raise ArchiveError("filename too long")
    block = n * ENTRY_FIXED + sum(len(b) for b in enc)
    out = bytearray(CDR_HEADER_SIZE + block)
    struct.pack_into("<IQQ", out, 0, SIG_CDR, n, block)
    p = CDR_HEADER_SIZE
    for i, b in enumerate(enc):
        struct.pack_into("<H", out, p, len(b))
        out[p + 2:p + 2 + len(b)] = b
        p += 2 + len(b)
        struct.pack_into("<QQQQB", out, p, int(offset[i]), int(comp[i]), int(uncomp[i]), int(hashes[i]), int(method[i]))
        p += 33
    return bytes(out)


def assemble(names: Sequence[str], payloads: Sequence, uncomp, hashes, method) -> np.ndarray:
    """Whole archive from already-compressed payloads: offsets = exclusive prefix sum from 10."""
    comp = np.array([len(p) for p in payloads], np.uint64)
    offset = DATA_START + np.concatenate([[0], np.cumsum(comp)[:-1]]).astype(np.uint64) if len(comp) else comp
    cdr_off = DATA_START + int(comp.sum())
    cdr = cdr_bytes(names, offset, comp, uncomp, hashes, method)
    total = cdr_off + len(cdr) + EOCDR_SIZE
    out = np.empty(total, np.uint8)
    out[:DATA_START] = np.frombuffer(struct.pack("<IHI", SIG_HEADER, 1, SIG_DATA), np.uint8)
    pos = DATA_START
    for p in payloads:
        k = len(p)
        out[pos:pos + k] = np.frombuffer(p, np.uint8) if not isinstance(p, np.ndarray) else p
        pos += k
    out[pos:pos + len(cdr)] = np.frombuffer(cdr, np.uint8)
    pos += len(cdr)
    out[pos:] = np.frombuffer(struct.pack("<IQ", SIG_EOCDR, cdr_off), np.uint8)
    return out
