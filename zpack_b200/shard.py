"""Static partition of an archive's entries across the GPUs of one box.

Entries are independent (own frame(s), own digest), so there is no data-path collective: every rank
gets a contiguous range of entries (in archive order) balanced by decoded bytes, reads only the byte
range of the archive that range touches, and writes only its own outputs.  The only cross-rank values
are O(entries) scalars the host already owns (comp_size[] for the offset prefix sum when packing,
status[] / digest[] when unpacking)."""
from __future__ import annotations

import numpy as np


def partition(uncomp_size, world: int):
    """-> list of (first, last_exclusive) per rank: contiguous, covering, balanced on sum(uncomp_size)."""
    u = np.asarray(uncomp_size, np.float64)
    n = len(u)
    if world <= 1 or n == 0:
        return [(0, n)] + [(n, n)] * max(0, world - 1)
    cum = np.cumsum(u)
    total = cum[-1] if n else 0.0
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(cum, target, side="left")) + 1 if total > 0 else (n * r) // world
        k = max(cuts[-1], min(n, k))
        # pick the nearer boundary
        if k > cuts[-1] + 1 and abs(cum[k - 2] - target) < abs(cum[k - 1] - target):
            k -= 1
        cuts.append(k)
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def byte_range(offset, comp_size, first: int, last: int):
    """[lo, hi) of the archive that entries [first, last) touch (what this rank copies to its GPU)."""
    if last <= first:
        return 0, 0
    off = np.asarray(offset[first:last], np.uint64)
    cs = np.asarray(comp_size[first:last], np.uint64)
    return int(off.min()), int((off + cs).max())


def pack_offsets(comp_size, data_start: int = 10):
    """entry.offset for GPU-packed entries: exclusive prefix sum of comp_size from the data start
    (header 6 + data signature 4; /root/reference/lib/zpack.h:41-42, lib/zpack_write.c:338)."""
    c = np.asarray(comp_size, np.uint64)
    return (np.uint64(data_start) + np.concatenate([[0], np.cumsum(c)[:-1]]).astype(np.uint64)) if len(c) else c


# ---------------------------------------------------------------------------------------------- C5
def split_blocks(n_blocks: int, world: int, block_size: int = 65536, total_size: int | None = None):
    """Intra-entry sharding of ONE block-independent LZ4 entry (BASELINE config C5):
    -> list of (first_block, last_block_exclusive) per rank, contiguous runs of near-equal length.
    Every block but the entry's last decodes to `block_size` bytes, so rank r's output starts at
    first_block * block_size.  The last NON-EMPTY shard keeps at least two blocks when the entry's last
    block is short (the XXH3 tail needs the final 64 bytes and the last partial KiB in one place)."""
    world = max(1, world)
    cuts = [(n_blocks * r) // world for r in range(world + 1)]
    if total_size is not None and n_blocks >= 2 and total_size - (n_blocks - 1) * block_size < 2048:
        for r in range(world, 0, -1):          # find the shard that holds the last block
            if cuts[r - 1] < cuts[r]:
                if cuts[r] - cuts[r - 1] < 2:  # a lone short last block: hand it to the previous shard's owner
                    cuts[r - 1] = max(0, cuts[r - 1] - 1)
                    for q in range(r - 2, 0, -1):
                        cuts[q] = min(cuts[q], cuts[q + 1])
                break
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def relay_digest(rank: int, world: int, chain, send, recv, empty: bool = False):
    """The only cross-GPU data of the sharded entry: the 64-byte XXH3 accumulator state, handed from
    the rank holding shard k to the rank holding shard k+1, in shard order (the scramble chain is
    serial, xxhash.h:3527-3534).  `chain(acc_in) -> (acc_out, digest_or_None)` runs this rank's part
    (zpb_blocks_digest); `recv(src) -> uint64[8] | None` and `send(dst, acc)` move the state.  Ranks
    with no blocks (`empty`) pass the state through.  Returns the digest on the rank that finished the
    entry, None elsewhere."""
    acc = recv(rank - 1) if rank > 0 else None
    digest = None
    if not empty:
        acc, digest = chain(acc)
    if rank + 1 < world:
        send(rank + 1, acc)
    return digest
