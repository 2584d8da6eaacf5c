#!/usr/bin/env python
"""Shared-memory wavefronts per CUDA source line from an ncu report captured with --import-source on.
usage: ncu_wavefronts.py report.ncu-rep [topN]"""
import csv, subprocess, sys, os
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = "?"
hdr = None
agg = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = os.path.basename(r[1]); continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == "-":           # per-line aggregate row
        iw = hdr.index("L1 Wavefronts Shared"); ie = hdr.index("L1 Wavefronts Shared Excessive"); ii = hdr.index("Instructions Executed")
        w = int(r[iw] or 0)
        if w:
            agg.append((w, int(r[ie] or 0), int(r[ii] or 0), cur, r[0], r[1].strip()[:105]))
tot = sum(a[0] for a in agg)
print("total shared wavefronts", tot, " excessive", sum(a[1] for a in agg))
for w, ex, ins, f, ln, src in sorted(agg, reverse=True)[:top]:
    print("%5.1f%%  exc %4.1f%%  %s:%s  %s" % (100.0 * w / tot, 100.0 * ex / tot, f, ln, src))
