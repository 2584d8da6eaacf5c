#!/usr/bin/env python
"""Generates tests/golden/large_golden.npz: outcomes of the numpy oracle (oracle/socp_oracle.py with block-structured
scaling, ``dense_scaling=False`` -- the same per-cone arithmetic as src/scalings.jl:32-99 without the k x k matrices,
LAPACK Cholesky / triangular solves as src/densesolver.jl:41-90 calls them) on BASELINE.json's large configurations,
which the GPU box cannot afford to run through the oracle at test time:

  C5  n=4096, k=8192, 64 x SOC(128), one problem:
        * the whole solve: status, iteration count, objectives, x
        * one step from seeded interior (s, z) and a seeded right-hand side: compute_scaling (lambda), setup_iter,
          solve_kkt -> cx, cz, cs  (pins src/densesolver.jl:41-90 at that size)
  C4  n=500, k=1000, 20 x SOC(50): 32 problems sampled from the 1000-problem batch (indices 0, 31, 62, ...):
        status, iteration count, objectives

The generators are counter-based (seed = 1234 + problem index), so the GPU test regenerates the same problems.

    python tests/golden/make_golden_large.py          # rewrites large_golden.npz (about 5 minutes on 8 cores)
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p_ in (ROOT, os.path.join(ROOT, "socp.jl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p_)
from oracle import socp_oracle as so          # noqa: E402
from socp_b200 import generators as gen       # noqa: E402

C4_SAMPLE = list(range(0, 1000, 31))[:32]


def step_inputs(cones, n, k, seed):
    """Seeded strictly interior (s, z) and a right-hand side (dx, dz, ds); shared with tests/test_gpu_fullsize.py."""
    rng = np.random.default_rng(seed)
    s, z = np.empty(k), np.empty(k)
    for kind, offs, dim in cones:
        for v in (s, z):
            if kind == 0:
                v[offs:offs + dim] = rng.uniform(0.5, 2.0, dim)
            else:
                tail = rng.standard_normal(dim - 1)
                v[offs + 1:offs + dim] = tail
                v[offs] = np.linalg.norm(tail) + rng.uniform(0.5, 1.5)
    return s, z, rng.standard_normal(n), rng.standard_normal(k), rng.standard_normal(k)


def main():
    out = {}
    t0 = time.time()
    prob = gen.make_config("C5")
    cones = tuple((c.kind, c.offs, c.dim) for c in prob.cones)
    pr = so.Problem.create(prob.c[0], prob.A_dense(0), prob.b[0], prob.G_dense(0), prob.h[0], cones, sing=False)
    print("C5 generated", time.time() - t0, flush=True)
    s, z, dx, dz, ds = step_inputs(cones, pr.n, pr.k, 77)
    sc = so.Scaling.create(cones, dense=False)
    so.compute_scaling(cones, sc, s, z)
    solver = so.DenseSolver(pr)
    solver.setup_iter(pr, sc)
    cx, cy, cz, cs = solver.solve_kkt(pr, sc, dx, np.zeros(0), dz, ds, fast_iprod=True)
    out.update(c5_step_lambda=sc.l.copy(), c5_step_cx=cx, c5_step_cz=cz, c5_step_cs=cs)
    print("C5 step done", time.time() - t0, flush=True)
    r = so.solve_socp(pr, init="reduced", fast_iprod=True, dense_scaling=False)
    out.update(c5_status=np.int32(r.status), c5_iters=np.int32(r.iters), c5_pobj=np.float64(r.pobj),
               c5_dobj=np.float64(r.dobj), c5_x=r.state.x)
    print("C5 solve done", r.status, r.iters, r.pobj, r.dobj, time.time() - t0, flush=True)

    st, it, po, do = [], [], [], []
    for q in C4_SAMPLE:
        p4 = gen.make_config("C4", batch=1, first=q)
        cones4 = tuple((c.kind, c.offs, c.dim) for c in p4.cones)
        pr4 = so.Problem.create(p4.c[0], p4.A_dense(0), p4.b[0], p4.G_dense(0), p4.h[0], cones4, sing=False)
        r4 = so.solve_socp(pr4, init="reduced", fast_iprod=True, dense_scaling=False)
        st.append(r4.status); it.append(r4.iters); po.append(r4.pobj); do.append(r4.dobj)
        print("C4", q, r4.status, r4.iters, r4.pobj, time.time() - t0, flush=True)
    out.update(c4_index=np.array(C4_SAMPLE, dtype=np.int32), c4_status=np.array(st, dtype=np.int32),
               c4_iters=np.array(it, dtype=np.int32), c4_pobj=np.array(po), c4_dobj=np.array(do))
    np.savez_compressed(os.path.join(HERE, "large_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "large_golden.npz"))


if __name__ == "__main__":
    main()
