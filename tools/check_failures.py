#!/usr/bin/env python
"""Solve a config on the GPU, then run the C oracle on the problems whose status is not CONVERGED
(and a few that are) and print both outcomes side by side."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "socp.jl_b200"))
import numpy as np
import socp_b200 as sb
from socp_b200 import generators as gen
from oracle import c_oracle as co
ap = argparse.ArgumentParser(); ap.add_argument("config"); ap.add_argument("--batch", type=int, default=None)
ap.add_argument("--path", default="auto"); ap.add_argument("--max-check", type=int, default=6)
a = ap.parse_args()
prob = gen.make_config(a.config, batch=a.batch)
ss = sb.SolverState(prob)
r = sb.solve_socp_batch(prob, ss, sb.default_params(path={"auto": 0, "tiled": 1, "fused": 2}[a.path]))
bad = np.nonzero(r.status != 0)[0]
print("status counts", np.bincount(r.status, minlength=3).tolist(), "bad idx", bad[:20].tolist())
good = np.nonzero(r.status == 0)[0][:2]
idx = np.concatenate([bad[:a.max_check], good]).astype(int)
cones = tuple((c.kind, c.offs, c.dim) for c in prob.cones)
t0 = time.time()
o = co.solve_batch(prob.c[idx], prob.A_cm[idx] if prob.p else np.zeros((len(idx), prob.n, 0)), prob.b[idx], prob.G_cm[idx], prob.h[idx], cones,
                   sing=np.zeros(len(idx), np.uint8), nthreads=os.cpu_count())
print("oracle time %.1f s" % (time.time() - t0))
for j, q in enumerate(idx):
    print(f"  prob {q:5d}: gpu status {r.status[q]} iters {r.iters[q]:2d} pobj {r.pobj[q]: .10e} dobj {r.dobj[q]: .10e} | "
          f"oracle status {o['status'][j]} iters {o['iters'][j]:2d} pobj {o['pobj'][j]: .10e} dobj {o['dobj'][j]: .10e}")
