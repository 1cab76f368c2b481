#!/usr/bin/env python
"""Opcode histogram (weighted by executed warp instructions) and hottest lines of an `ncu --page source --csv` dump."""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
ia, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
data = []
for r in rows[hi + 1:]:
    try:
        data.append((int(r[ia]), int(r[isamp]), r[isrc].strip()))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
ts = sum(d[1] for d in data)
print("total warp-inst", tot, "samples", ts)
c, s = Counter(), Counter()
for n, sm, src in data:
    t = src.split()
    op = t[1] if t[0].startswith("@") else t[0]
    op = op.split(".")[0]
    c[op] += n
    s[op] += sm
for op, n in c.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print("%-10s %10d %5.1f%%  samples %5d %5.1f%%" % (op, n, 100 * n / tot, s[op], 100 * s[op] / max(ts, 1)))
