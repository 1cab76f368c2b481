#!/usr/bin/env python
"""Tabulate a tools/trace_layer.py dump: one row per tile, one column per (role) event."""
import re, sys
ev = {}
for line in open(sys.argv[1]):
    m = re.match(r"\s*(\d+)\s+w(\d+)\s+tile\s+(\d+)\s+(\S+)", line)
    if not m:
        continue
    t, w, tile, name = int(m[1]), int(m[2]), int(m[3]), m[4]
    if w in (0, 1, 2, 18, 26) or (len(sys.argv) > 3 and w == int(sys.argv[3])):
        ev.setdefault(tile, {})[name] = t
cols = sys.argv[2].split(",") if len(sys.argv) > 2 else sorted({k for v in ev.values() for k in v})
print("tile " + " ".join("%10s" % c.split(":")[1][:10] for c in cols))
for tile in sorted(ev):
    print("%4d " % tile + " ".join("%10s" % ev[tile].get(c, "-") for c in cols))
