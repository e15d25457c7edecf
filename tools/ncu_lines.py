#!/usr/bin/env python
"""Per-source-line summary of an `ncu --page source --csv --print-source cuda,sass` dump (instructions executed,
stall samples, average active lanes).  Usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | ncu_lines.py"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
cur = None
agg = []
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        ci, si, ti = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
        continue
    if hdr and len(r) > ti and r[0].isdigit() and r[2] == "-":      # a source line (SASS rows carry an address)
        try:
            agg.append((int(r[ci]), int(r[si]), int(r[ti]), cur, int(r[0]), r[1].strip()[:100]))
        except ValueError:
            pass
tot = sum(a[0] for a in agg) or 1
tots = sum(a[1] for a in agg) or 1
print("total warp instructions %d, stall samples %d" % (tot, tots))
top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for a in sorted(agg, reverse=True)[:top]:
    print("%5.1f%% inst %5.1f%% samples lanes %4.1f | %s:%d | %s" % (100 * a[0] / tot, 100 * a[1] / tots, a[2] / max(a[0], 1), a[3], a[4], a[5]))
