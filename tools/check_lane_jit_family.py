#!/usr/bin/env python
"""Randomly drawn layouts of the lane kernel's family (orthant rows, then second-order cones of dimension 2..8, p = 0,
n <= 16), each specialised at run time and compared with fused_v2's one-warp teams on a small batch.
usage: python tools/check_lane_jit_family.py [count] [seed]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "socp.jl_b200"))
import numpy as np
import socp_b200 as sb
from socp_b200 import generators as gen
count = int(sys.argv[1]) if len(sys.argv) > 1 else 12
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
for t in range(count):
    n = int(rng.integers(2, 17))
    kpoc = int(rng.integers(0, 6))
    cones, at = ([sb.POC(0, kpoc)] if kpoc else []), kpoc
    for _ in range(int(rng.integers(1, 5))):
        d, c = int(rng.integers(2, 9)), int(rng.integers(1, 5))
        for _ in range(c):
            cones.append(sb.SOC(at, d)); at += d
    if at < n or at > 96:
        continue
    prob = gen.random_feasible(700, n, 0, cones, 0.3, 0, 100 + t)
    out = {}
    for lane in ("1", "0"):
        os.environ["SOCP_B200_LANE"] = lane
        ss = sb.SolverState(prob); ss.load(prob)
        out[lane] = sb.solve_socp_batch(prob, ss, reload=False)
    a, b = out["1"], out["0"]
    conv = (a.status == 0) & (b.status == 0) & (a.iters == b.iters)
    d = np.abs(a.pobj[conv] - b.pobj[conv]) / np.maximum(1.0, np.abs(b.pobj[conv]))
    st = (a.status == b.status).mean()
    ok = st >= 0.98 and (d.size == 0 or np.quantile(d, 0.99) < 1e-6)
    bad += not ok
    print(f"{'ok ' if ok else 'BAD'} n={n:2d} k={at:2d} kpoc={kpoc} cones={[c.dim for c in cones if c.kind == 1]} status equal {st:.4f} "
          f"converged {(a.status == 0).mean():.3f} same iters {conv.mean():.3f} obj diff max {d.max() if d.size else 0:.1e}")
print("layouts with disagreement:", bad)
