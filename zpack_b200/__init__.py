"""zpack_b200 — B200-native (sm_100a) implementation of ZPack's per-entry hot path.

Only what the path needs lives here:
  csrc/            CUDA kernels + the C-ABI of include/zpack_b200.h  -> libzpack_b200.so
  host/            C++ mirror of the reference's lib/zpack.h API      -> libzpack.so (drop-in)
  lib.py           ctypes binding of the C-ABI (tests / bench use the product through it)
  container.py     ZPack container framing (docs/specs.md) for assembling / parsing archives
  corpus.py        the deterministic synthetic corpus "zpk-synth-v1" 

There is no CPU codec in this package: without a CUDA device `lib.Context()` raises.
"""
from .lib import Context, Group, ZpbError, Entry, File, ArcEntry, load_library  # noqa: F401

__all__ = ["Context", "Group", "ZpbError", "Entry", "File", "ArcEntry", "load_library"]
