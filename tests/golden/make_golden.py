#!/usr/bin/env python
"""Generates tests/golden/solve_golden.json: outcomes of the numpy oracle (oracle/socp_oracle.py, the restatement
pinned to the reference's own golden vectors) on (a) the literal instances of the reference's test file
(tests/refcases.py = /root/reference/test/runtests.jl:130-244) and (b) small seeded batches of the BASELINE.json
shapes.  The reference is Julia and cannot run in this image, so these fixtures freeze the
ORACLE's answers: status, iteration count, objectives and the iterate x, per problem.

    python tests/golden/make_golden.py          # rewrites solve_golden.json
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p_ in (ROOT, os.path.join(ROOT, "socp.jl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p_)
from oracle import socp_oracle as so          # noqa: E402
from socp_b200 import generators as gen       # noqa: E402
import refcases as rc                         # noqa: E402

SEEDED = {"C2": 6, "C3": 12}                  # config -> problems (gen.make_config(name, batch=...), first = 0)


def solve_one(c, A, b, G, h, cones, sing=None, faithful=False):
    """faithful: the reference's own initial system (src/solver.jl:68-84) and O(d^2) iprod loop (src/vectors.jl:105-125);
    otherwise the block-eliminated initial point and the O(d) closed form, as the batch tests use."""
    pr = so.Problem.create(c, A, b, G, h, cones, sing=sing)
    r = so.solve_socp(pr, init="full", fast_iprod=False) if faithful else so.solve_socp(pr, init="reduced", fast_iprod=True)
    return dict(status=int(r.status), iters=int(r.iters), pobj=float(r.pobj), dobj=float(r.dobj),
                x=[float(v) for v in r.state.x])


def main():
    out = {"reference_instances": {}, "seeded": {}}
    for name, make in rc.ALL_C1.items():
        d = make()
        out["reference_instances"][name] = solve_one(d["c"], d["A"], d["b"], d["G"], d["h"], d["cones"], faithful=True)
    for cfg, B in SEEDED.items():
        prob = gen.make_config(cfg, batch=B)
        cones = tuple((c.kind, c.offs, c.dim) for c in prob.cones)
        out["seeded"][cfg] = [solve_one(prob.c[q], prob.A_dense(q), prob.b[q], prob.G_dense(q), prob.h[q], cones, sing=False)
                              for q in range(B)]
    with open(os.path.join(HERE, "solve_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", os.path.join(HERE, "solve_golden.json"))


if __name__ == "__main__":
    main()
