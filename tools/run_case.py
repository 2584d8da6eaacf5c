#!/usr/bin/env python
"""Runs one configuration once (for ncu / sanitizer captures): python tools/run_case.py C2 --batch 592 --path fused"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "socp.jl_b200"))
import numpy as np  # noqa: E402
import socp_b200 as sb  # noqa: E402
from socp_b200 import generators as gen  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("config")
ap.add_argument("--batch", type=int, default=None)
ap.add_argument("--path", default="auto", choices=["auto", "tiled", "fused"])
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--max-iter", type=int, default=40)
a = ap.parse_args()
prob = gen.make_config(a.config, batch=a.batch)
ss = sb.SolverState(prob)
ss.load(prob)
prm = sb.default_params(path={"auto": 0, "tiled": 1, "fused": 2}[a.path], max_iter=a.max_iter)
for _ in range(a.reps):
    r = sb.solve_socp_batch(prob, ss, prm, reload=False, want_iterates=False)
print(a.config, "batch", prob.B, "status counts", np.bincount(r.status, minlength=3).tolist(), "mean iters %.2f" % r.iters.mean(),
      "timings", r.timings)
