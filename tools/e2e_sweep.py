#!/usr/bin/env python
"""Developer tool: sweep the pipelined host path's chunk size / worker count on the C2 archive
(zpb_unpack_host with pinned buffers; full round trip and ZPB_F_DISCARD).  One JSON line per setting."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    import zpack_b200
    from zpack_b200 import container
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    arch = bench.build_archive(n, bench.ENTRY_SIZE)
    d = container.parse(arch)
    e = d.entries()
    out_size = int(e["dst_off"][-1] + e["dst_cap"][-1])
    h_arch = torch.from_numpy(arch).pin_memory().numpy()
    h_out = torch.empty(out_size, dtype=torch.uint8).pin_memory().numpy()
    ev = e.copy()
    ev["flags"] |= 2
    unc = float(d.uncomp_size.sum())
    for chunk in [int(x) for x in os.environ.get("SWEEP_CHUNKS", "128,256,512,1024,2048").split(",")]:
        for workers in [int(x) for x in os.environ.get("SWEEP_WORKERS", "2,3,4,6").split(",")]:
            os.environ["ZPB_HOST_CHUNK_MB"], os.environ["ZPB_HOST_WORKERS"] = str(chunk), str(workers)
            ctx = zpack_b200.Context(0)
            res = {}
            for name, ent in (("full", e), ("verify", ev)):
                ctx.unpack_host(h_arch, len(arch), h_out, out_size, ent)
                ts = []
                for _ in range(3):
                    t0 = time.perf_counter()
                    st, dg = ctx.unpack_host(h_arch, len(arch), h_out, out_size, ent)
                    ts.append(time.perf_counter() - t0)
                assert (st == 0).all() and np.array_equal(dg, d.hash)
                res[name] = round(unc / min(ts) / 1e9, 1)
            print(json.dumps({"chunk_mb": chunk, "workers": workers, "GBps": res}), flush=True)
            ctx.close()


if __name__ == "__main__":
    main()
