#!/usr/bin/env python
"""Developer tool: how fast does the GPU reader decode the frames the GPU zstd writer produces (many small sub-blocks,
treeless literals, repeated tables) next to frames written by the reference (one block per entry)?"""
import json
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import zpack_b200
    from zpack_b200 import lib as zlib, corpus
    n, size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 131072
    ctx = zpack_b200.Context(0)
    data = np.concatenate([corpus.entry_bytes(i, size) for i in range(n)])
    cap = ctx.pack_bound(1, size)
    slot = (cap + 15) & ~15
    f = np.zeros(n, zlib.File)
    f["src_off"] = np.arange(n, dtype=np.uint64) * size
    f["size"], f["dst_off"], f["dst_cap"], f["method"] = size, np.arange(n, dtype=np.uint64) * slot, cap, 1
    d_in = torch.from_numpy(data).cuda()
    d_out = torch.empty(n * slot, dtype=torch.uint8, device="cuda")
    comp, dg, st = ctx.pack_device(d_in, d_in.numel(), d_out, d_out.numel(), f)
    assert (st == 0).all()
    e = np.zeros(n, zlib.Entry)
    e["src_off"], e["comp_size"], e["uncomp_size"] = f["dst_off"], comp, size
    e["dst_off"] = np.arange(n, dtype=np.uint64) * size
    e["dst_cap"], e["hash"], e["method"] = size, dg, 1
    d_back = torch.empty(n * size, dtype=torch.uint8, device="cuda")
    ms = []
    for r in range(4):
        st2, dg2 = ctx.unpack_device(d_out, d_out.numel(), d_back, d_back.numel(), e)
        ms.append(ctx.last_kernel_ms()["unpack_ms"])
    assert (st2 == 0).all() and torch.equal(d_back, d_in)
    t = float(np.median(ms[1:]))
    print(json.dumps({"entries": n, "ratio": round(n * size / float(comp.sum()), 3), "gpu_written_zstd_unpack_GBps": round(n * size / t / 1e6, 1), "ms": round(t, 3)}))


if __name__ == "__main__":
    main()
