#!/usr/bin/env python
"""Static code size (SASS instructions) and, with an .ncu-rep, executed instructions / stall samples per OUTERMOST
source line of a kernel: nvdisasm -gi gives the inlining chain of every instruction; the last entry is the line of
the kernel body that the instruction was inlined into.  Shows which call sites make the kernel big (instruction-cache
footprint) and where the samples fall.
usage: sass_by_callsite.py object.o kernel_substring [report.ncu-rep] [topN]"""
import collections, csv, os, re, subprocess, sys, tempfile
obj, kname = sys.argv[1], sys.argv[2]
rep = sys.argv[3] if len(sys.argv) > 3 and sys.argv[3].endswith(".ncu-rep") else None
top = int(sys.argv[-1]) if sys.argv[-1].isdigit() else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
instrs = []            # (outermost (file, line), innermost (file, line), text)
infn, chain = False, []
for l in sass:
    if l.startswith("//--------------------- .text."):
        infn = kname in l
        continue
    if not infn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        chain.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        if chain:
            last = (chain[-1], chain[0])
        instrs.append((last[0], last[1], m.group(2)))
        chain = []
data = None
if rep:
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    hdr, data = None, []
    for r in csv.reader(out.splitlines()):
        if len(r) > 3 and r[0] == "Address":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
for i, (outer, inner, txt) in enumerate(instrs):
    a = agg[outer]
    a[0] += 1
    if data and i < len(data):
        a[1] += int(data[i]["Instructions Executed"] or 0)
        a[2] += int(data[i]["# Samples"] or 0)
        for k, v in data[i].items():
            if k.startswith("stall_") and "Not Issued" not in k and v and v != "0":
                a[3][k[6:]] += int(v)
srcs = {}
def src(f, ln):
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "socp.jl_b200", "csrc", f)
    if os.path.exists(p):
        if p not in srcs:
            srcs[p] = open(p).read().splitlines()
        return srcs[p][ln - 1].strip()[:90] if ln - 1 < len(srcs[p]) else ""
    return ""
n = len(instrs)
te = sum(a[1] for a in agg.values()) or 1
ts = sum(a[2] for a in agg.values()) or 1
print(f"{n} SASS instructions ({n * 16 / 1024:.0f} KB)")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ",".join(f"{k}:{100 * v // max(1, sum(a[3].values()))}%" for k, v in a[3].most_common(2))
    print(f"{a[0]:5d} sass {100 * a[0] / n:4.1f}%  exec {100 * a[1] / te:4.1f}%  smp {100 * a[2] / ts:4.1f}% [{st}]  {f}:{ln}  {src(f, ln)}")
