#!/usr/bin/env python
"""Rank CUDA source lines of an `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` dump by stall samples
and executed warp instructions."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
hdr, curfile, out = None, "", []
for r in rows:
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        curfile = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 10 or r[0] == "" or r[0] == "Function Name":
        continue
    si, ii, ti = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    try:
        out.append((curfile, r[0], r[1].strip()[:105], float(r[si] or 0), float(r[ii] or 0), float(r[ti] or 0)))
    except ValueError:
        pass
ts = sum(o[3] for o in out)
tins = sum(o[4] for o in out)
print("total samples %.0f, warp instructions %.3e" % (ts, tins))
for o in sorted(out, key=lambda o: -o[3])[:top]:
    print("%-13s %4s | samp %5.2f%% inst %5.2f%% act %4.1f | %s" % (o[0][:13], o[1], 100 * o[3] / max(ts, 1), 100 * o[4] / max(tins, 1), o[5] / max(o[4], 1), o[2]))
