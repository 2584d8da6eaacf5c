#!/usr/bin/env python
"""Device-resident solve time of a tiny layout WITHOUT a compile-time instantiation of the lane-per-problem kernel: the
kernel specialised at run time (NVRTC, csrc/lane_jit.cu) against fused_v2's one-warp teams.
usage: python tools/bench_lane_jit.py [batch]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "socp.jl_b200"))
import numpy as np
import socp_b200 as sb
from socp_b200 import generators as gen
B = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
layouts = {"n=8: POC 2 + 6 x SOC(3)": (8, [sb.POC(0, 2)] + [sb.SOC(2 + 3 * i, 3) for i in range(6)]),
           "n=10: 8 x SOC(5)": (10, [sb.SOC(5 * i, 5) for i in range(8)]),
           "n=9: POC 3 + 4 x SOC(3) + 3 x SOC(5) + SOC(4)": (9, [sb.POC(0, 3)] + [sb.SOC(3 + 3 * i, 3) for i in range(4)] +
                                                             [sb.SOC(15 + 5 * i, 5) for i in range(3)] + [sb.SOC(30, 4)]),
           "n=14: 12 x SOC(4)": (14, [sb.SOC(4 * i, 4) for i in range(12)]),
           "n=16: POC 8 + 8 x SOC(4)": (16, [sb.POC(0, 8)] + [sb.SOC(8 + 4 * i, 4) for i in range(8)])}
if len(sys.argv) > 2:
    layouts = {k: v for k, v in layouts.items() if sys.argv[2] in k}
for name, (n, cones) in layouts.items():
    prob = gen.random_feasible(B, n, 0, cones, 0.3, 0, 21)
    keep = {}
    for lane in ("1", "0"):
        os.environ["SOCP_B200_LANE"] = lane
        ss = sb.SolverState(prob)
        t0 = time.time(); ss.load(prob); sb.solve_socp_batch(prob, ss, reload=False); first = time.time() - t0
        ms = min(sb.solve_socp_batch(prob, ss, reload=False).timings["solve_ms"] for _ in range(3))
        res = sb.solve_socp_batch(prob, ss, reload=False)
        keep[lane] = res
        print(f"{name:28s} batch {B} {'lane kernel (run-time specialised)' if lane == '1' else 'one-warp teams (fused_v2)          '}: "
              f"{ms:8.3f} ms = {B / ms / 1e3:6.2f}M problems/s, converged {(res.status == 0).mean():.4f}, first call {first:.1f} s")
    a, b = keep["1"], keep["0"]
    same = (a.status == b.status) & (a.iters == b.iters) & (a.status == 0)
    d = np.abs(a.pobj[same] - b.pobj[same]) / np.maximum(1.0, np.abs(b.pobj[same]))
    print(f"    status equal {(a.status == b.status).mean():.4f}, same iterations {same.mean():.4f}, objective difference max {d.max():.2e} median {np.median(d):.1e}")
