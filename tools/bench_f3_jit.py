#!/usr/bin/env python
"""Device-resident solve time of C2-like layouts other than BASELINE.json's own (portfolio SOCPs with n != 50): the
whole-solve kernel of fused_v3.cuh specialised for the layout at run time (NVRTC, csrc/lane_jit.cu) against its
runtime-dimension instantiation.  usage: python tools/bench_f3_jit.py [batch]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "socp.jl_b200"))
import numpy as np
import socp_b200 as sb
from socp_b200 import generators as gen
B = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
for n in (40, 30, 64):
    prob = gen.portfolio(B, n)
    keep = {}
    for jit in (True, False):
        if jit: os.environ.pop("SOCP_B200_NO_F3_JIT", None)
        else: os.environ["SOCP_B200_NO_F3_JIT"] = "1"
        ss = sb.SolverState(prob)
        t0 = time.time(); ss.load(prob); sb.solve_socp_batch(prob, ss, reload=False); first = time.time() - t0
        ms = min(sb.solve_socp_batch(prob, ss, reload=False).timings["solve_ms"] for _ in range(3))
        keep[jit] = sb.solve_socp_batch(prob, ss, reload=False)
        print(f"portfolio n={n:2d} batch {B} {'specialised at run time  ' if jit else 'runtime-dimension kernel'}: {ms:8.3f} ms = "
              f"{B / ms:8.1f}k problems/s, converged {(keep[jit].status == 0).mean():.4f}, first call {first:.1f} s")
    a, b = keep[True], keep[False]
    same = (a.status == b.status) & (a.iters == b.iters) & (a.status == 0)
    d = np.abs(a.pobj[same] - b.pobj[same]) / np.maximum(1.0, np.abs(b.pobj[same]))
    print(f"    status equal {(a.status == b.status).mean():.4f}, same iterations {same.mean():.4f}, objective difference max {d.max():.2e}")
