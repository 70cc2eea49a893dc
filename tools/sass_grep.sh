#!/bin/bash
# per kernel of the shipped library: how many bulk-async copy (UBLKCP), mbarrier (SYNCS) and per-thread async copy
# (LDGSTS) instructions its SASS holds.  Usage: tools/sass_grep.sh > profiles/r2_sass_bulk_copy.txt
echo "# cuobjdump -sass zpack_b200/libzpack_b200.so (sm_100a): count, kernel, instruction"
cuobjdump -sass zpack_b200/libzpack_b200.so | python3 -c '
import re, sys, collections
c = collections.Counter(); fn = "?"
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m: fn = m.group(1); continue
    m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?((?:UBLKCP|SYNCS|LDGSTS)[A-Z0-9_.]*)", line)
    if m: c[(fn, m.group(1))] += 1
for (fn, op), n in sorted(c.items()): print(n, fn, op)
' | c++filt | sed -E 's/\(.*\)//'
