#!/usr/bin/env python
"""Device-resident solve time of a tiny layout with equality rows (outside the lane kernel's family): fused_v2's one-warp
teams specialised for the layout at run time (NVRTC) against the runtime-dimension instantiation.
usage: python tools/bench_f2_jit.py [batch]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "socp.jl_b200"))
import numpy as np
import socp_b200 as sb
from socp_b200 import generators as gen
B = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
for name, n, p, cones in (("n=10 p=2: POC 4 + 6 x SOC(3)", 10, 2, [sb.POC(0, 4)] + [sb.SOC(4 + 3 * i, 3) for i in range(6)]),
                          ("n=12 p=3: 10 x SOC(4)", 12, 3, [sb.SOC(4 * i, 4) for i in range(10)])):
    prob = gen.random_feasible(B, n, p, cones, 0.3, 0, 31)
    keep = {}
    for jit in (True, False):
        if jit: os.environ.pop("SOCP_B200_NO_F2_JIT", None)
        else: os.environ["SOCP_B200_NO_F2_JIT"] = "1"
        ss = sb.SolverState(prob); ss.load(prob); sb.solve_socp_batch(prob, ss, reload=False)
        ms = min(sb.solve_socp_batch(prob, ss, reload=False).timings["solve_ms"] for _ in range(3))
        keep[jit] = sb.solve_socp_batch(prob, ss, reload=False)
        print(f"{name:30s} batch {B} {'specialised at run time  ' if jit else 'runtime-dimension kernel'}: {ms:8.3f} ms = "
              f"{B / ms / 1e3:6.2f}M problems/s, converged {(keep[jit].status == 0).mean():.4f}")
    a, b = keep[True], keep[False]
    same = (a.status == b.status) & (a.iters == b.iters) & (a.status == 0)
    d = np.abs(a.pobj[same] - b.pobj[same]) / np.maximum(1.0, np.abs(b.pobj[same]))
    print(f"    status equal {(a.status == b.status).mean():.4f}, same iterations {same.mean():.4f}, objective difference max {d.max():.2e}")
