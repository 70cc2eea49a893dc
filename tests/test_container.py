"""Host-side container framing (zpack_b200/container.py) against the reference's golden archives."""
import os

import numpy as np
import pytest

from zpack_b200 import container


@pytest.mark.parametrize("kind,size", [("none", 648), ("zstd", 241), ("lz4", 270)])  # tests/archive.h:99-101
def test_parse_golden(golden_dir, kind, size):
    raw = np.fromfile(os.path.join(golden_dir, f"archive_{kind}.zpk"), np.uint8)
    assert len(raw) == size
    d = container.parse(raw)
    assert d.names == ["file1.txt", "file2.txt"]
    assert list(d.uncomp_size) == [169, 349]
    assert list(d.hash) == [0x7874CBA47D02B07D, 0x15F25C0F24DD8E52]
    assert list(d.method) == [{"none": 0, "zstd": 1, "lz4": 2}[kind]] * 2
    assert int(d.offset[0]) == 10


def test_assemble_reproduces_golden_none_archive(golden_dir):
    raw = np.fromfile(os.path.join(golden_dir, "archive_none.zpk"), np.uint8)
    files = [open(os.path.join(golden_dir, n), "rb").read() for n in ("file1.txt", "file2.txt")]
    out = container.assemble(["file1.txt", "file2.txt"], files, [169, 349],
                             [0x7874CBA47D02B07D, 0x15F25C0F24DD8E52], [0, 0])
    assert np.array_equal(out, raw)


def test_reject_malformed():
    with pytest.raises(container.ArchiveError):
        container.parse(np.zeros(10, np.uint8))
    with pytest.raises(container.ArchiveError):
        container.parse(np.zeros(64, np.uint8))


def test_empty_archive_roundtrip():
    out = container.assemble([], [], [], [], [])
    assert len(out) == container.MIN_ARCHIVE
    assert len(container.parse(out)) == 0


def test_entry_table_layout():
    out = container.assemble(["a", "b"], [b"xyz", b"12345"], [3, 5], [1, 2], [0, 0])
    e = container.parse(out).entries()
    assert list(e["src_off"]) == [10, 13] and list(e["dst_off"]) == [0, 16] and list(e["dst_cap"]) == [3, 5]
