"""CPU-side checks of the drop-in boundary: the CUDA library loads and exports every symbol that
include/zpack_b200.h declares (no compute calls — there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from zpack_b200 import lib as zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zp(?:b|ack)_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(zlib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(zlib.LIB_PATH)
    names = _declared("zpack_b200.h")
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/zpack_b200.h but not exported"
    assert set(zlib.EXPORTS) <= set(names)
    assert lib.zpb_abi_version() == 1


def test_descriptor_layouts_are_64_bytes():
    assert zlib.Entry.itemsize == 64 and zlib.File.itemsize == 64
    assert zlib.Entry.fields["hash"][1] == 40 and zlib.Entry.fields["method"][1] == 48


def test_header_structs_match_the_python_records(tmp_path):
    """include/zpack_b200.h compiled as plain C: size and field offsets of every record equal the numpy dtypes of the binding."""
    import subprocess
    recs = {"zpb_entry": zlib.Entry, "zpb_file": zlib.File, "zpb_block": zlib.Block, "zpb_arc_entry": zlib.ArcEntry}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "zpack_b200.h"', 'int main(void) {']
    for name, dt in recs.items():
        lines.append(f'printf("{name} __sizeof %zu\\n", sizeof({name}));')
        for f in dt.names:
            lines.append(f'printf("{name} {f} %zu\\n", offsetof({name}, {f}));')
    lines += ['return 0; }']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    got = {tuple(l.split()[:2]): int(l.split()[2]) for l in out if l}
    for name, dt in recs.items():
        assert got[(name, "__sizeof")] == dt.itemsize, name
        for f in dt.names:
            assert got[(name, f)] == dt.fields[f][1], (name, f)


def test_no_gpu_means_loud_failure():
    """No CPU fallback: without a device the context constructor raises instead of degrading."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(zlib.ZpbError):
        zlib.Context(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "zpack_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".inl")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower(), f"{os.path.join(dirpath, f)} mentions the oracle"
