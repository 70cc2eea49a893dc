#!/usr/bin/env python
"""Hottest instructions (by stall samples) of a captured kernel with their dominant stall reasons.
  python tools/ncu_hot.py report.ncu-rep [min_samples]"""
import csv, io, subprocess, sys
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; col = {h: i for i, h in enumerate(hdr)}
sc = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 1500
base = None; tot = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    tot += float(r[col['# Samples']] or 0)
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    a = int(r[0], 16)
    if base is None: base = a
    smp = float(r[col['# Samples']] or 0)
    if smp > thr:
        st = sorted(((float(r[col[c]] or 0), c) for c in sc), reverse=True)[:2]
        print(f"{a-base:06x} {100*smp/tot:5.2f}% inst {float(r[col['Instructions Executed']] or 0):.3e} {r[col['Source']][:58]:58s} {[(c.replace('stall_',''), int(v)) for v, c in st]}")
