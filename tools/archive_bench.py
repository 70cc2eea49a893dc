"""Device-resident container operations on a C2-shaped archive (65 536 entries, ~57 KB of payload each): archive build from
pack slots (offset table + directory kernels, payload copy kernel), entry copy into a second archive, directory parse.
One JSON line per measurement; kernel times are CUDA events recorded by the library around its own launches
(zpb_last_archive_ms).  Run under gpurun:  python tools/archive_bench.py [entries]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import zpack_b200  # noqa: E402
from zpack_b200 import container  # noqa: E402
from zpack_b200.lib import ArcEntry  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6456.5))
    ctx = zpack_b200.Context(0)
    rng = np.random.default_rng(1)
    comp = rng.integers(30000, 84000, n).astype(np.uint64)              # C2's compressed sizes: mean ~57 KB
    names = [f"dir{i % 64:02d}/entry_{i:07d}.bin" for i in range(n)]
    blob = np.frombuffer("".join(names).encode(), np.uint8)
    e = np.zeros(n, ArcEntry)
    e["comp_size"], e["uncomp_size"], e["hash"], e["method"] = comp, 131072, rng.integers(0, 2**63, n).astype(np.uint64), 2
    e["name_len"] = [len(s) for s in names]
    e["name_off"] = np.concatenate([[0], np.cumsum(e["name_len"])[:-1]])
    for label, slot_align in (("slots 16-byte aligned (pack output)", 16), ("slots at arbitrary offsets", 1)):
        cap = (comp + np.uint64(131072 // 4)).astype(np.uint64)        # slot capacity > payload, as after a pack
        if slot_align == 16:
            cap = (cap + np.uint64(15)) & ~np.uint64(15)
        else:
            cap = cap | np.uint64(1)
        e["src_off"] = np.concatenate([[0], np.cumsum(cap)[:-1]]) + (0 if slot_align == 16 else 3)
        src_size = int(cap.sum()) + 16
        d_src = torch.randint(0, 256, (src_size,), dtype=torch.uint8, device="cuda")
        total = 10 + int(comp.sum()) + 20 + 35 * n + len(blob) + 12
        d_arch = torch.empty(total + 64, dtype=torch.uint8, device="cuda")
        best = None
        for it in range(6):
            t0 = time.perf_counter()
            size = ctx.archive_build_device(d_src, src_size, e, blob, d_arch, total + 64)
            wall = (time.perf_counter() - t0) * 1e3
            ms = ctx.last_archive_ms()
            if it >= 2 and (best is None or ms[1] < best[1]):
                best = (ms[0], ms[1], wall)
        moved = 2 * int(comp.sum())
        print(json.dumps({"op": "archive_build_device", "case": label, "entries": n, "payload_GB": round(int(comp.sum()) / 1e9, 3),
                          "table_and_directory_ms": round(best[0], 4), "copy_ms": round(best[1], 4), "call_wall_ms": round(best[2], 3),
                          "copy_GBps_read_plus_write": round(moved / best[1] / 1e6, 1), "frac_of_hbm_peak": round(moved / best[1] / 1e6 / hbm, 3),
                          "hbm_peak_GBps": hbm}), flush=True)
        if slot_align == 1:
            # parse the directory just written; then copy every second entry into another archive (zpack_write_files_from_archive)
            for it in range(4):
                t0 = time.perf_counter()
                res, eo, nb = ctx.archive_open_device(d_arch, size)
                wall = (time.perf_counter() - t0) * 1e3
            assert res == 0 and np.array_equal(eo["hash"], e["hash"]) and np.array_equal(eo["offset"], e["offset"])
            t0 = time.perf_counter()
            d = container.parse(d_arch[:size].cpu().numpy())
            host_ms = (time.perf_counter() - t0) * 1e3
            print(json.dumps({"op": "archive_open_device", "entries": n, "directory_MB": round(len(nb) / 1e6, 2),
                              "kernels_ms": round(ctx.last_archive_ms()[2], 4), "call_wall_ms_incl_table_download": round(wall, 3),
                              "python_host_parse_ms_incl_archive_download": round(host_ms, 1)}), flush=True)
            sub = np.ascontiguousarray(eo[::2])
            d_new = torch.empty(total + 64, dtype=torch.uint8, device="cuda")
            for it in range(4):
                ctx.archive_build_device(d_arch, size, sub, nb, d_new, total + 64)
            ms = ctx.last_archive_ms()
            moved = 2 * int(sub["comp_size"].sum())
            print(json.dumps({"op": "archive to archive (every second entry)", "entries": len(sub), "copy_ms": round(ms[1], 4),
                              "copy_GBps_read_plus_write": round(moved / ms[1] / 1e6, 1), "frac_of_hbm_peak": round(moved / ms[1] / 1e6 / hbm, 3)}),
                  flush=True)
        del d_src, d_arch
    ctx.close()


if __name__ == "__main__":
    main()
