#!/usr/bin/env python
"""Annotated SASS listing (executed count, stall samples, top stall reasons) of the instructions that map to a
source line range.  usage: ncu_sass.py report.ncu-rep lib.so kernel_substring file lo hi"""
import csv, os, re, subprocess, sys, tempfile
rep, so, kname, fname, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
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
acc = 0
for i in range(min(len(lines), len(data))):
    (f, ln), ins = lines[i]; d = data[i]
    if f == fname and lo <= ln <= hi:
        sp = int(d["# Samples"] or 0); acc += sp
        st = sorted(((int(d[c] or 0), c[6:]) for c in stallcols), reverse=True)[:2]
        sts = ",".join(f"{n}:{v}" for v, n in st if v)
        print(f"{i:6d} L{ln:<4d} ex={d['Instructions Executed']:>8s} smp={sp:5d} [{sts:28s}] {ins[:90]}")
print("samples in range", acc, "of", tot, f"({100*acc/tot:.1f}%)")
