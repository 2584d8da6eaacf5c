#!/usr/bin/env python
"""Splits a kernel's SASS at BAR.SYNC instructions and reports, per segment (in SASS order): warp-instructions
executed, stall samples by reason and the source lines seen -- a barrier-to-barrier latency profile.
usage: ncu_phases.py report.ncu-rep lib.so kernel_substring [min_samples]"""
import csv, os, re, subprocess, sys, tempfile, collections
rep, so, kname = sys.argv[1], sys.argv[2], sys.argv[3]
minsmp = int(sys.argv[4]) if len(sys.argv) > 4 else 50
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines = []; infn = False; cur = ("?", 0)
for l in sass:
    if l.startswith("//--------------------- .text."):
        infn = kname in l; continue
    if not infn: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: lines.append((cur, m.group(2)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr = None; data = []
for r in rows:
    if len(r) > 3 and r[0] == "Address": hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(dict(zip(hdr, r)))
stallcols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(d["# Samples"] or 0) for d in data)
seg = dict(start=0, inst=0, smp=0, st=collections.Counter(), src=collections.Counter(), n=0)
def flush(end):
    if seg["smp"] >= minsmp:
        st = ",".join(f"{k}:{v}" for k, v in seg["st"].most_common(3))
        src = ",".join(f"{f.split('.')[0][-8:]}:{l}" for (f, l), _ in seg["src"].most_common(4))
        print(f"[{seg['start']:5d}-{end:5d}] sass={seg['n']:4d} inst={seg['inst']:9d} smp={seg['smp']:6d} {100*seg['smp']/tot:5.1f}% [{st}] {src}")
for i in range(min(len(lines), len(data))):
    (f, ln), ins = lines[i]; d = data[i]
    sp = int(d["# Samples"] or 0)
    seg["inst"] += int(d["Instructions Executed"] or 0); seg["smp"] += sp; seg["n"] += 1
    seg["src"][(f, ln)] += sp
    for c in stallcols:
        v = int(d[c] or 0)
        if v: seg["st"][c[6:]] += v
    if ins.startswith("BAR.SYNC") or "WARPSYNC" in ins and False:
        flush(i)
        seg = dict(start=i + 1, inst=0, smp=0, st=collections.Counter(), src=collections.Counter(), n=0)
flush(len(lines))
print("total samples", tot)
