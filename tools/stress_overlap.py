"""Developer stress test of the overlapped parse / execute launch: many back-to-back calls on mixed archives of
different sizes, every call checked (status + digest of every entry, decoded bytes of a sample)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import zpack_b200
from zpack_b200 import container, corpus
from class_bench import build

ctx = zpack_b200.Context(0)
bad = 0
for n, reps in ((37, 300), (1000, 200), (4096, 150), (16384, 100)):
    arch = build(-1, n, 131072, False, 2)
    d = container.parse(arch); e = d.entries()
    out_size = int(e["dst_off"][-1] + e["dst_cap"][-1])
    d_arch = torch.from_numpy(arch).cuda(); d_out = torch.empty(out_size, dtype=torch.uint8, device="cuda")
    want = torch.from_numpy(corpus.entry_bytes(1, 131072)).cuda()
    for r in range(reps):
        d_out.zero_()
        st, dg = ctx.unpack_device(d_arch, len(arch), d_out, out_size, e)
        if not ((st == 0).all() and np.array_equal(dg, d.hash)):
            bad += 1
        o = int(e["dst_off"][1])
        if not torch.equal(d_out[o:o + 131072], want):
            bad += 1
    print(f"n={n}: {reps} calls, failures so far {bad}", flush=True)
    del d_arch, d_out
print("FAILURES", bad)
sys.exit(1 if bad else 0)
