#!/usr/bin/env python
"""Joins an ncu `--page source --csv` SASS dump with nvdisasm line info and aggregates
executed instructions / stall samples per CUDA source line.
usage: ncu_by_line.py report.ncu-rep lib.so kernel_substring [topN]"""
import csv, os, re, subprocess, sys, tempfile, collections
rep, so, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
sass = []
for f in sorted(os.listdir(tmp)):          # one cubin per translation unit: take the one that holds the kernel
    if f.endswith(".cubin"):
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if kname in txt:
            sass = txt.splitlines()
            break
# instruction list (in order) of the wanted kernel with (file, line)
lines = []
infn = False
cur = ("?", 0)
for l in sass:
    if l.startswith("//--------------------- .text."):
        infn = kname in l
        continue
    if not infn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        lines.append((cur, m.group(2)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
data = []
for r in rows:
    if len(r) > 3 and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(dict(zip(hdr, r)))
print("sass instrs:", len(lines), "ncu rows:", len(data))
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
stallcols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
n = min(len(lines), len(data))
for i in range(n):
    key = lines[i][0]
    d = data[i]
    a = agg[key]
    a[0] += int(d["Instructions Executed"] or 0)
    a[1] += int(d["# Samples"] or 0)
    for sc in stallcols:
        v = int(d[sc] or 0)
        if v:
            a[2][sc] += v
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print("total warp-inst", ti, "samples", ts)
srcs = {}
def src(f, ln):
    for root in ("/root/repo/socp.jl_b200/csrc",):
        p = os.path.join(root, f)
        if os.path.exists(p):
            if p not in srcs: srcs[p] = open(p).read().splitlines()
            if 0 < ln <= len(srcs[p]): return srcs[p][ln - 1].strip()[:90]
    return ""
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    st = ", ".join(f"{k_[6:]}:{v}" for k_, v in a[2].most_common(3))
    print(f"{100*a[1]/max(ts,1):5.1f}% smp {100*a[0]/max(ti,1):5.1f}% inst  {key[0]}:{key[1]:<4d} [{st}]  {src(*key)}")

# ---- buckets by function (line ranges given on the command line via env BUCKETS="name:file:lo-hi,...")
bk = os.environ.get("BUCKETS")
if bk:
    print("\nbuckets:")
    rest_i, rest_s = ti, ts
    for spec in bk.split(","):
        name, f, rng = spec.split(":")
        lo, hi = [int(v) for v in rng.split("-")]
        bi = sum(a[0] for (ff, ln), a in agg.items() if ff == f and lo <= ln <= hi)
        bs = sum(a[1] for (ff, ln), a in agg.items() if ff == f and lo <= ln <= hi)
        rest_i -= bi; rest_s -= bs
        print(f"  {name:14s} {100*bs/max(ts,1):5.1f}% smp {100*bi/max(ti,1):5.1f}% inst")
    print(f"  {'(rest)':14s} {100*rest_s/max(ts,1):5.1f}% smp {100*rest_i/max(ti,1):5.1f}% inst")
