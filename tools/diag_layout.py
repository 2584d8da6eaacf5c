#!/usr/bin/env python
"""Per-problem comparison fused / tiled / C oracle / numpy oracle on one generic layout (debugging aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "socp.jl_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import socp_b200 as sb
from socp_b200 import generators as gen
from oracle import c_oracle as co, socp_oracle as so
L = {"mixed_p2": (20, 2, (sb.POC(0, 6), sb.SOC(6, 9), sb.SOC(15, 5))),
     "mixed_n60_p9": (60, 9, (sb.POC(0, 10), sb.SOC(10, 30), sb.SOC(40, 30), sb.SOC(70, 30)))}
for name in sys.argv[1:]:
    n, p, cones = L[name]
    prob = gen.random_feasible(24, n, p, cones, 0.1)
    oc = tuple((c.kind, c.offs, c.dim) for c in prob.cones)
    f = sb.solve_socp_batch(prob, sb.SolverState(prob))
    t = sb.solve_socp_batch(prob, sb.SolverState(prob), sb.default_params(path=sb.PATH_TILED))
    r = co.solve_batch(prob.c, prob.A_cm, prob.b, prob.G_cm, prob.h, oc, sing=prob.sing, nthreads=4)
    print(name)
    for q in range(24):
        pr = so.Problem.create(prob.c[q], prob.A_dense(q), prob.b[q], prob.G_dense(q), prob.h[q], oc, sing=False)
        o = so.solve_socp(pr, init="reduced", fast_iprod=True)
        print(f" {q:2d} fused {f.status[q]} {f.iters[q]:2d} {f.pobj[q]:+.10f} | tiled {t.status[q]} {t.iters[q]:2d} {t.pobj[q]:+.10f} |"
              f" C {r['status'][q]} {r['iters'][q]:2d} {r['pobj'][q]:+.10f} | numpy {o.status} {o.iters:2d} {o.pobj:+.10f}")
