#!/usr/bin/env python
"""Per-source-line instruction / stall-sample listing of one file's line range from an ncu report.
usage: ncu_lines.py report.ncu-rep lib.so kernel_substring file lo hi"""
import csv, os, re, subprocess, sys, tempfile, collections
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
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
stallcols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ti = ts = 0
for i in range(min(len(lines), len(data))):
    (f, ln), ins = lines[i]; d = data[i]
    ie = int(d["Instructions Executed"] or 0); sp = int(d["# Samples"] or 0)
    ti += ie; ts += sp
    if f == fname and lo <= ln <= hi:
        a = agg[ln]; a[0] += ie; a[1] += sp; a[2] += 1
        for sc in stallcols:
            v = int(d[sc] or 0)
            if v: a[3][sc[6:]] += v
src = open(os.path.join("/root/repo/socp.jl_b200/csrc", fname)).read().splitlines()
print("total inst", ti, "samples", ts)
for ln in sorted(agg):
    a = agg[ln]
    st = ",".join(f"{k}:{v}" for k, v in a[3].most_common(2))
    print(f"{ln:4d} {100*a[1]/ts:5.2f}%smp {100*a[0]/ti:5.2f}%inst sass={a[2]:4d} [{st}] {src[ln-1].strip()[:80]}")
