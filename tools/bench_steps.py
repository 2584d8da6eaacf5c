#!/usr/bin/env python
"""Per-kernel roofline numbers of the step-level (tiled-path) kernels: compute_scaling / scale! / vprod! / iprod! /
compute_step against the HBM peak, SYRK / Cholesky against the FP64 peak (SURVEY.md section 8(d) byte and flop counts).
usage: python tools/bench_steps.py C3 [--batch N] [--reps R] [--only 0,7,8]"""
import argparse, ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "socp.jl_b200"))
import numpy as np
import socp_b200 as sb
from socp_b200 import generators as gen, _lib as L
ap = argparse.ArgumentParser(); ap.add_argument("config"); ap.add_argument("--batch", type=int, default=None)
ap.add_argument("--reps", type=int, default=10); ap.add_argument("--only", default=None)
a = ap.parse_args()
prob = gen.make_config(a.config, batch=a.batch)
B, n, p, k, N = prob.B, prob.n, prob.p, prob.k, len(prob.cones)
ss = sb.SolverState(prob); ss.load(prob)
rng = np.random.default_rng(1)
# a strictly interior (s, z)
s = np.empty((B, k)); z = np.empty((B, k))
for c in prob.cones:
    sl = slice(c.offs, c.offs + c.dim)
    for v in (s, z):
        if c.kind == 0: v[:, sl] = rng.uniform(0.5, 2.0, (B, c.dim))
        else:
            t = rng.standard_normal((B, c.dim - 1)); v[:, c.offs + 1:c.offs + c.dim] = t
            v[:, c.offs] = np.linalg.norm(t, axis=1) + rng.uniform(0.5, 1.5, B)
sb.compute_scaling(prob.cones, ss.scaling, s, z)
hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
fp64 = float(os.environ.get("FP64_PEAK_TFLOPS", "35.3"))     # cuBLAS DGEMM measured by bench.py on this pool
spec = {  # which: (name, bytes per problem, flops per problem)   SURVEY.md section 8(d)
    0: ("compute_scaling", 32 * k + 8 * N, 0), 1: ("scale!", 24 * k, 0), 2: ("iscale!", 24 * k, 0), 3: ("vprod!", 24 * k, 0),
    4: ("iprod!", 24 * k, 0), 5: ("compute_step", 24 * k, 0), 6: ("Gt = W^-1 G", 16 * k * n, 0),
    7: ("SYRK Gt'Gt", 8 * k * n + 4 * n * (n + 1), n * (n + 1) * k), 8: ("Cholesky", 8 * n * n, n ** 3 / 3.0),
    9: ("L L' solve (1 rhs)", 8 * n * n, 2 * n * n), 10: ("G'v", 8 * k * n, 2 * k * n), 11: ("G v", 8 * k * n, 2 * k * n)}
lib = L.load()
out = C.c_double()
sel = [int(v) for v in a.only.split(",")] if a.only else sorted(spec)
print(f"{a.config}: batch {B}, n={n}, p={p}, k={k}, {N} cones; peaks: HBM {hbm:.0f} GB/s (measured), FP64 {fp64:.1f} TFLOP/s (measured DGEMM)")
for w in sel:
    name, byt, fl = spec[w]
    ss.handle.check(lib.socp_b200_profile_step(ss.handle.ptr, w, a.reps, C.byref(out)), "profile_step")
    ms = out.value
    gbs = byt * B / (ms * 1e-3) / 1e9
    line = f"  {w:2d} {name:20s} {ms*1e3:10.1f} us   {gbs:8.1f} GB/s = {100*gbs/hbm:5.1f}% of HBM"
    if fl: line += f"   {fl*B/(ms*1e-3)/1e12:7.2f} TFLOP/s = {100*fl*B/(ms*1e-3)/1e12/fp64:5.1f}% of FP64"
    print(line)
