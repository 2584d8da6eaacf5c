#!/usr/bin/env python
"""Small end-to-end solves for compute-sanitizer runs (racecheck / memcheck): the fused kernel on the two specialised
shapes, on generic layouts with and without equality rows, and the tiled path on a mid-size layout."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "socp.jl_b200"))
import numpy as np
import socp_b200 as sb
from socp_b200 import generators as gen
cases = [("C2", gen.make_config("C2", batch=3)), ("C3", gen.make_config("C3", batch=5)),
         ("mixed p=2", gen.random_feasible(3, 20, 2, (sb.POC(0, 6), sb.SOC(6, 9), sb.SOC(15, 5)), 0.1)),
         ("n=40 p=0", gen.random_feasible(2, 40, 0, gen.soc_cones(4, 12), 0.1))]
if "--lane" in sys.argv:      # the lane-per-problem kernel (run with SOCP_B200_LANE=1): several problems per lane
    cases = [("C3 lane", gen.make_config("C3", batch=300)),
             ("orthant + cones lane", gen.random_feasible(40, 6, 0, [sb.POC(0, 5)] + [sb.SOC(5 + 3 * i, 3) for i in range(3)], 0.3, 0, 11))]
if "--tiled" in sys.argv:
    cases = [("tiled n=150", gen.random_feasible(2, 150, 3, gen.soc_cones(6, 30), 0.05))]
for name, prob in cases:
    res = sb.solve_socp_batch(prob, sb.SolverState(prob))
    print(name, "path", res.timings["path_used"], "status", res.status.tolist(), "iters", res.iters.tolist())
