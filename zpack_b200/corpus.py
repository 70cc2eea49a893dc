"""Synthetic corpus "zpk-synth-v1" (SURVEY.md appendix G): deterministic from (seed, entry index).

Entry i of size E has class i mod 4:
  0 random   uniform bytes                        -> LZ4 stored blocks / zstd raw blocks
  1 text     Zipf-distributed words from a fixed 4096-word vocabulary, space separated
  2 runs     (byte, length in [1,512]) runs       -> offset-1 overlapping matches, long lengths
  3 records  64-byte records: LE32 counter, 4 random bytes at +8, 56 constant bytes per entry

The CPU generator is the single source of truth; the GPU only ever consumes the bytes.
"""
from __future__ import annotations

import numpy as np

SEED = 0x5A504B15
CLASSES = ("random", "text", "runs", "records")
_VOCAB = None


def _rng(seed: int, index: int, stream: int = 0) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([seed, index, stream]))


def _vocab():
    """4096 lowercase words of length 2..10 (+ trailing space), as a padded byte matrix."""
    global _VOCAB
    if _VOCAB is None:
        g = np.random.Generator(np.random.PCG64([SEED, 0x766F6361]))
        lens = g.integers(2, 11, size=4096)
        mat = g.integers(ord("a"), ord("z") + 1, size=(4096, 11), dtype=np.uint8)
        col = np.arange(11)[None, :]
        mat[col == lens[:, None]] = ord(" ")
        flat = mat[col <= lens[:, None]]                      # words + trailing space, concatenated
        wlen = (lens + 1).astype(np.int64)
        wstart = np.cumsum(wlen) - wlen
        p = 1.0 / np.arange(1, 4097)
        cdf = np.cumsum(p / p.sum())
        # inverse CDF quantised to 16 bits: rank-frequency ~ 1/rank, one table lookup per word
        table = np.searchsorted(cdf, (np.arange(65536) + 0.5) / 65536.0).clip(0, 4095).astype(np.int64)
        _VOCAB = (flat, wstart, wlen, table)
    return _VOCAB


def entry_bytes(index: int, size: int, seed: int = SEED) -> np.ndarray:
    """The `size` bytes of entry `index` (uint8 array)."""
    cls = index & 3
    g = _rng(seed, index)
    if size == 0:
        return np.zeros(0, np.uint8)
    if cls == 0:
        return np.frombuffer(g.bytes(size), np.uint8).copy()
    if cls == 1:
        flat, wstart, wlen, table = _vocab()
        nwords = size // 6 + 16
        while True:
            idx = table[g.integers(0, 65536, size=nwords)]
            L = wlen[idx]
            if L.sum() >= size:
                break
            nwords *= 2
        ends = np.cumsum(L)
        src = np.repeat(wstart[idx] - (ends - L), L) + np.arange(ends[-1])
        return np.ascontiguousarray(flat[src[:size]])
    if cls == 2:
        nruns = size // 128 + 16
        while True:
            lens = g.integers(1, 513, size=nruns)
            if lens.sum() >= size:
                break
            nruns *= 2
        vals = g.integers(0, 256, size=nruns, dtype=np.uint8)
        return np.ascontiguousarray(np.repeat(vals, lens)[:size])
    nrec = (size + 63) // 64
    rec = np.empty((nrec, 64), np.uint8)
    rec[:] = g.integers(0, 256, size=64, dtype=np.uint8)[None, :]
    rec[:, 0:4] = np.arange(nrec, dtype="<u4").view(np.uint8).reshape(nrec, 4)
    rec[:, 8:12] = g.integers(0, 256, size=(nrec, 4), dtype=np.uint8)
    return np.ascontiguousarray(rec.reshape(-1)[:size])


def entry_name(index: int) -> str:
    return f"{CLASSES[index & 3]}/{index:08d}.bin"


def generate(n_entries: int, size: int, first: int = 0, seed: int = SEED) -> np.ndarray:
    """(n_entries, size) uint8 matrix of consecutive entries."""
    out = np.empty((n_entries, size), np.uint8)
    for k in range(n_entries):
        out[k] = entry_bytes(first + k, size, seed)
    return out


def big_entry_piece(k: int, piece: int, total: int, first: int = 0, seed: int = SEED) -> np.ndarray:
    """Piece k of the single large entry of BASELINE config C5: the four classes cycling every
    `piece` bytes (1 MiB in the bench), truncated at `total`."""
    lo = k * piece
    return entry_bytes(first + k, max(0, min(piece, total - lo)), seed)


def big_entry(total: int, piece: int = 1 << 20, first: int = 0, seed: int = SEED) -> np.ndarray:
    n = (total + piece - 1) // piece
    return np.concatenate([big_entry_piece(k, piece, total, first, seed) for k in range(n)]) if n else np.zeros(0, np.uint8)
