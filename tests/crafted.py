"""Hand-built LZ4 frames aimed at specific branches of the execute kernel (shared by the GPU test and the CPU-emulation
test).  The expected bytes come from the builder itself (a direct restatement of the sequence semantics,
lz4_Block_format.md); callers cross-check them with the oracle's decoder."""
import numpy as np

LINKED_HDR = bytes([0x04, 0x22, 0x4D, 0x18, 0x40, 0x40, 0xC0])   # what zpack_write_files emits (tests/golden)
OFFSETS = [1, 2, 3, 4, 5, 7, 8, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 511, 512, 513, 1023, 1024, 2048, 4000,
           4095, 4096, 4097, 5000, 8191, 8192, 20000, 40000, 65535]


def emit(comp, lit_bytes, off, ml):
    lit = len(lit_bytes)
    mlc = ml - 4 if ml else 0
    comp.append((min(lit, 15) << 4) | min(mlc, 15))
    if lit >= 15:
        r = lit - 15
        comp.extend(b"\xff" * (r // 255))
        comp.append(r % 255)
    comp.extend(lit_bytes)
    if ml:
        comp.extend(int(off).to_bytes(2, "little"))
        if mlc >= 15:
            r = mlc - 15
            comp.extend(b"\xff" * (r // 255))
            comp.append(r % 255)


def build_block(rng, history: bytearray, size: int, style: str):
    """Appends `size` decoded bytes to history; returns the compressed block."""
    comp = bytearray()
    start = len(history)
    end = start + size
    while True:
        room = end - len(history)
        if room <= 32:
            break
        if style == "chains":        # fixed-stride records: every match sources the previous one
            lit, off, ml = int(rng.integers(0, 5)), 64, int(rng.choice([7, 20, 52, 53, 60]))
        elif style == "fills":
            lit, off, ml = int(rng.integers(0, 3)), int(rng.choice([1, 2, 4])), int(rng.choice([4, 31, 32, 33, 64, 200, 1000, 5000]))
        elif style == "long":
            lit, off, ml = int(rng.choice([0, 1, 17, 33, 70, 300])), int(rng.choice(OFFSETS)), int(rng.choice([32, 48, 64, 65, 100, 300, 700, 2000, 6000]))
        else:
            lit = int(rng.choice([0, 0, 0, 1, 2, 3, 8, 15, 16, 17, 31, 32, 33, 40, 255, 270]))
            off = int(rng.choice(OFFSETS))
            ml = int(rng.choice([4, 5, 8, 12, 16, 17, 18, 19, 20, 33, 64, 65, 66, 274]))
        if lit + ml + 16 > room:
            break
        lit_bytes = rng.integers(0, 256, size=lit, dtype=np.uint8).tobytes()
        avail = len(history) + lit
        if off > avail or (avail - off < 0):
            off = max(1, avail) if avail else 0
        if avail == 0:
            lit_bytes = rng.integers(0, 256, size=8, dtype=np.uint8).tobytes()
            history.extend(lit_bytes)
            emit(comp, lit_bytes, 1, 4)
            history.extend(history[-1:] * 4)
            continue
        off = min(off, avail, 65535)
        history.extend(lit_bytes)
        src = len(history) - off
        if off >= ml:
            history.extend(history[src:src + ml])
        else:
            pat = bytes(history[src:])
            history.extend((pat * (ml // off + 1))[:ml])
        emit(comp, lit_bytes, off, ml)
    tail = rng.integers(0, 256, size=end - len(history), dtype=np.uint8).tobytes()   # >= 12 literals end the block
    history.extend(tail)
    emit(comp, tail, 0, 0)
    assert len(history) == end
    return bytes(comp)


def frame(rng, sizes, style):
    history = bytearray()
    fr = bytearray(LINKED_HDR)
    for s in sizes:
        blk = build_block(rng, history, s, style)
        assert len(blk) <= 65536 * 2
        fr += len(blk).to_bytes(4, "little") + blk
    fr += b"\0\0\0\0"
    return bytes(fr), bytes(history)


