"""CPU model of how the execute kernel (zpack_b200/csrc/lz4_fast.cuh, K2) orders the 32 sequences of one step.

K2 copies all literals of a step first, then every match whose source bytes are already final — in parallel, in any
order — and only then the few matches it could not resolve, in sequence order.  A match whose source lies wholly
inside an earlier match of the same step is *redirected* through that match's offset (out[x] = out[x - off] for every
x of the parent's match region), shifts being composed by pointer jumping.  This file restates exactly that rule in
Python (the classification, the composition and the execution order — not the CUDA data movement, which the GPU
tests cover) and checks on real LZ4 streams and on hand-built ones that executing the redirected matches against a
SNAPSHOT taken before any match of the step was written reproduces the plaintext.  If the rule were wrong for some
shape of dependency (parent overlapping itself, source straddling two sequences, chains deeper than the jumping
rounds), the snapshot execution would read stale bytes and the comparison would fail here, without a GPU."""
import struct

import numpy as np
import pytest

from zpack_b200 import corpus

FINAL, CHILD, HARD = 0, 1, 2


def sequences(block: bytes):
    """(out position, literal start in the block, literal length, offset, match length) per sequence."""
    p, n, op, out = 0, len(block), 0, []
    while p < n:
        t = block[p]
        p += 1
        lit = t >> 4
        if lit == 15:
            while True:
                b = block[p]
                p += 1
                lit += b
                if b != 255:
                    break
        lsrc = p
        p += lit
        o = op
        op += lit
        if p >= n:
            out.append((o, lsrc, lit, 0, 0))
            break
        off = block[p] | (block[p + 1] << 8)
        p += 2
        ml = t & 15
        if ml == 15:
            while True:
                b = block[p]
                p += 1
                ml += b
                if b != 255:
                    break
        ml += 4
        out.append((o, lsrc, lit, off, ml))
        op += ml
    return out


def classify(step):
    """The dependency pass of K2 for one step: state, shift per sequence (lz4_fast.cuh, `same-step dependencies`)."""
    n = len(step)
    o_first = step[0][0]
    state, shift, parent = [FINAL] * n, [0] * n, [0] * n
    for i, (o, _, lit, off, ml) in enumerate(step):
        if ml == 0:
            continue
        mo = o + lit
        msrc = mo - off
        send = min(msrc + ml, mo)
        if not (send > o_first and msrc < o):
            continue                                   # before the step, or inside this sequence's own literals
        k = 0                                          # largest sequence starting at or below msrc (5 shuffle rounds)
        for st in (16, 8, 4, 2, 1):
            if k + st < n and step[k + st][0] <= msrc:
                k += st
        ok_, _, lk, offk, mlk = step[k]
        mo_k, e_k = ok_ + lk, ok_ + lk + mlk
        if msrc < o_first:
            state[i] = HARD
        elif send <= mo_k:
            state[i] = FINAL                           # inside sequence k's literals
        elif msrc >= mo_k and send <= e_k and offk >= mlk and off >= ml:
            state[i], parent[i], shift[i] = CHILD, k, offk
        else:
            state[i] = HARD
    rounds = 0
    while any(s == CHILD for s in state):
        ps, psh, pp = list(state), list(shift), list(parent)
        for i in range(n):
            if state[i] == CHILD:
                k = parent[i]
                if ps[k] == HARD:
                    state[i], shift[i] = HARD, 0
                else:
                    shift[i] += psh[k]
                    parent[i] = pp[k]
                    if ps[k] == FINAL:
                        state[i] = FINAL
        rounds += 1
        assert rounds <= 6, "pointer jumping did not converge in log2(32) + 1 rounds"
    return state, shift


def execute_block(block: bytes, history: bytearray):
    """Appends the block's decoded bytes to `history`, step by step, the way K2 orders the work."""
    base = len(history)
    seqs = sequences(block)
    total = seqs[-1][0] + seqs[-1][2] + seqs[-1][4]
    history.extend(b"\0" * total)
    stats = [0, 0, 0]
    for s0 in range(0, len(seqs), 32):
        step = seqs[s0:s0 + 32]
        for (o, lsrc, lit, _, _) in step:              # literal phase
            history[base + o:base + o + lit] = block[lsrc:lsrc + lit]
        state, shift = classify(step)
        snap = bytes(history)                          # nothing of this step's matches exists yet
        for i, (o, _, lit, off, ml) in enumerate(step):
            if ml == 0 or state[i] != FINAL:
                continue
            mo = base + o + lit
            if shift[i] == 0 and off < ml:             # periodic match of final bytes: its own lane writes in order
                for j in range(ml):
                    history[mo + j] = history[mo + j - off]
            else:
                src = mo - off - shift[i]
                assert src >= 0 and src + ml <= mo
                history[mo:mo + ml] = snap[src:src + ml]
            stats[0 if shift[i] == 0 else 1] += 1
        for i, (o, _, lit, off, ml) in enumerate(step):   # what is left, in sequence order
            if ml == 0 or state[i] != HARD:
                continue
            mo = base + o + lit
            for j in range(ml):
                history[mo + j] = history[mo + j - off]
            stats[2] += 1
    return stats


def decode_frame_modelled(frame: bytes):
    pos, out, stats = 7, bytearray(), [0, 0, 0]
    linked = not (frame[4] >> 5) & 1
    while True:
        bh = struct.unpack_from("<I", frame, pos)[0]
        pos += 4
        if bh == 0:
            break
        bsz = bh & 0x7FFFFFFF
        blk = frame[pos:pos + bsz]
        pos += bsz
        if bh >> 31:
            out += blk
            continue
        if linked:
            st = execute_block(blk, out)
        else:
            part = bytearray()
            st = execute_block(blk, part)
            out += part
        stats = [a + b for a, b in zip(stats, st)]
    return bytes(out), stats


@pytest.mark.parametrize("cls", [1, 2, 3])
@pytest.mark.parametrize("independent", [False, True])
def test_step_model_on_the_corpus_classes(oracle, cls, independent):
    data = corpus.entry_bytes(cls, 131072 + 777)
    frame = bytes(oracle.lz4f_encode_port(data, 0, independent=independent))
    got, stats = decode_frame_modelled(frame)
    assert got == bytes(data)
    if cls in (1, 3):
        assert stats[1] > 0, "the corpus class is expected to exercise redirection"
        assert stats[2] < 0.1 * sum(stats), "only a small share of matches should be left for the ordered phase"


def test_step_model_on_hand_built_chains():
    """Chains deeper than one jump, a self-overlapping parent (children must NOT be redirected through it), a source
    that straddles two sequences, a source that starts before the step and ends inside it."""
    rng = np.random.default_rng(11)

    def emit(comp, lit_bytes, off, ml):
        lit, mlc = len(lit_bytes), (ml - 4 if ml else 0)
        comp.append((min(lit, 15) << 4) | min(mlc, 15))
        if lit >= 15:
            comp.extend(b"\xff" * ((lit - 15) // 255) + bytes([(lit - 15) % 255]))
        comp.extend(lit_bytes)
        if ml:
            comp.extend(int(off).to_bytes(2, "little"))
            if mlc >= 15:
                comp.extend(b"\xff" * ((mlc - 15) // 255) + bytes([(mlc - 15) % 255]))

    for trial in range(40):
        comp, plain = bytearray(), bytearray()
        first = rng.integers(0, 256, 200, dtype=np.uint8).tobytes()
        emit(comp, first, 100, 20)
        plain += first
        plain += plain[-100:-80]
        for _ in range(int(rng.integers(40, 200))):
            kind = int(rng.integers(0, 5))
            lit = rng.integers(0, 256, int(rng.choice([0, 0, 1, 2, 5])), dtype=np.uint8).tobytes()
            plain += lit
            if kind == 0:      # chain: source is the previous match
                off, ml = int(rng.choice([8, 16, 24, 64])), int(rng.choice([4, 7, 8, 16]))
            elif kind == 1:    # self-overlapping (periodic) match, later ones may point into it
                off, ml = int(rng.choice([1, 2, 3, 5])), int(rng.choice([9, 20, 40]))
            elif kind == 2:    # straddles the previous sequence boundary
                off, ml = int(rng.choice([5, 6, 7, 9])), int(rng.choice([4, 6, 8]))
            elif kind == 3:    # far back
                off, ml = int(rng.integers(100, len(plain))), int(rng.choice([4, 12, 33, 70]))
            else:
                off, ml = int(rng.integers(1, min(len(plain), 40) + 1)), int(rng.integers(4, 30))
            off = min(off, len(plain))
            src = len(plain) - off
            for j in range(ml):
                plain.append(plain[src + j])
            emit(comp, lit, off, ml)
        tail = rng.integers(0, 256, 12, dtype=np.uint8).tobytes()
        emit(comp, tail, 0, 0)
        plain += tail
        out = bytearray()
        execute_block(bytes(comp), out)
        assert bytes(out) == bytes(plain), trial
