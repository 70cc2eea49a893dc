#!/usr/bin/env python
"""Aggregate an ncu report's per-SASS-instruction counters by CUDA source line (no GUI needed).

  python tools/ncu_lines.py gpurun_out/prof.ncu-rep <kernel-regex> [--so zpack_b200/libzpack_b200.so] [--top 40]

Joins `ncu --page source --csv` (SASS rows, in program order) with `nvdisasm -g` line info of the
same kernel in the .so that was profiled, by instruction offset.  Inlined library headers are
attributed to the last repo line seen before them.
"""
import argparse
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(so, kernel_re):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    out = {}
    for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
        txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
        cur, line, inside = None, None, False
        for ln in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                inside = re.search(kernel_re, m.group(1)) is not None
                cur = m.group(1)
                line = None
                continue
            if not inside:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                if "/csrc/" in m.group(1):
                    line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                out[int(m.group(1), 16)] = (line, m.group(2).strip())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel")
    ap.add_argument("--so", default="zpack_b200/libzpack_b200.so")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--sass", action="store_true", help="also list the hottest SASS instructions")
    ap.add_argument("--dump", default=None, help="write the whole kernel, in address order, with per-instruction counts")
    ap.add_argument("--sass-re", default=None, help="regex on the MANGLED name in the .so (default: the kernel regex)")
    a = ap.parse_args()
    lines = sass_lines(a.so, a.sass_re or a.kernel)
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv", "--kernel-name", "regex:" + a.kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {h: i for i, h in enumerate(hdr)}
    base = None
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])  # inst, samples, shared wavefronts, thread inst
    hot = []
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr) or not r[0]:
            continue
        try:
            addr = int(r[0], 16) if r[0].startswith("0x") else int(r[0])
        except ValueError:
            continue
        if base is None:
            base = addr
        off = addr - base
        ln, txt = lines.get(off, (None, r[col["Source"]]))

        def num(k):
            try:
                return float(r[col[k]] or 0)
            except (KeyError, ValueError):
                return 0.0
        inst, smp = num("Instructions Executed"), num("# Samples")
        wf = num("L1 Wavefronts Shared")
        ti = num("Thread Instructions Executed")
        e = agg[ln]
        e[0] += inst; e[1] += smp; e[2] += wf; e[3] += ti
        hot.append((inst, smp, off, ln, r[col["Source"]]))
    tot_i = sum(v[0] for v in agg.values()) or 1
    tot_s = sum(v[1] for v in agg.values()) or 1
    print(f"total warp instructions {tot_i:.3e}, samples {tot_s:.0f}")
    print(f"{'line':>24} {'inst%':>7} {'smp%':>7} {'inst':>12} {'thr/inst':>8} {'smem wf':>12}")
    for ln, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:a.top]:
        name = f"{ln[0]}:{ln[1]}" if ln else "?"
        print(f"{name:>24} {100 * v[0] / tot_i:7.2f} {100 * v[1] / tot_s:7.2f} {v[0]:12.3e} {v[3] / max(v[0], 1):8.1f} {v[2]:12.3e}")
    if a.dump:
        with open(a.dump, "w") as f:
            for inst, smp, off, ln, txt in sorted(hot, key=lambda t: t[2]):
                name = f"{ln[0]}:{ln[1]}" if ln else "?"
                f.write(f"{off:06x} {name:>20} {inst:11.4e} {100 * smp / tot_s:5.2f}%  {txt}\n")
    if a.sass:
        print("\nhottest SASS (by samples):")
        for inst, smp, off, ln, txt in sorted(hot, key=lambda t: -t[1])[:a.top]:
            name = f"{ln[0]}:{ln[1]}" if ln else "?"
            print(f"  {off:06x} {name:>22} smp {100 * smp / tot_s:5.2f}% inst {inst:10.3e}  {txt}")


if __name__ == "__main__":
    main()
