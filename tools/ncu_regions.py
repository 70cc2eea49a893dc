#!/usr/bin/env python
"""Buckets the per-instruction dump of tools/ncu_lines.py (--dump) into regions of the execute kernel's body:
instructions of inlined helpers are attributed to the body statement that precedes them in address order.
  python tools/ncu_regions.py dump.txt 'name:lo-hi,name:lo-hi,...' [body_first_line]"""
import re, sys, collections
dump, spec = sys.argv[1], sys.argv[2]
first = int(sys.argv[3]) if len(sys.argv) > 3 else 0
regs = []
for part in spec.split(','):
    name, rng = part.split(':'); lo, hi = rng.split('-'); regs.append((name, int(lo), int(hi)))
agg = collections.OrderedDict((r[0], [0.0, 0.0]) for r in regs)
agg['other'] = [0.0, 0.0]
cur = 'other'
for ln in open(dump):
    m = re.match(r'([0-9a-f]{6})\s+(\S+)\s+([0-9.e+]+)\s+([0-9.]+)%\s+(.*)', ln)
    if not m: continue
    src, inst, smp = m.group(2), float(m.group(3)), float(m.group(4))
    if src.startswith('lz4_fast.cuh:'):
        l = int(src.split(':')[1])
        if l >= first:
            for name, lo, hi in regs:
                if lo <= l <= hi: cur = name; break
    agg[cur][0] += inst; agg[cur][1] += smp
tot = sum(v[0] for v in agg.values())
for k, v in agg.items():
    print(f"{k:>16} inst {v[0]:.3e} ({100*v[0]/tot:5.1f}%)  samples {v[1]:5.1f}%")
print(f"{'total':>16} inst {tot:.3e}")
