#!/usr/bin/env python
"""Per-phase cycle breakdown of the fused whole-solve kernel (CTA 0), using the -DSOCP_PHASE_TIMING build.
usage: python tools/phase_timing.py C2 --batch 2960"""
import argparse, ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "socp.jl_b200"))
from socp_b200 import build as B
os.environ["SOCP_B200_LIB"] = B.PROF_LIB_PATH
if not os.path.exists(B.PROF_LIB_PATH):
    B.build_prof()
import numpy as np
import socp_b200 as sb
from socp_b200 import generators as gen, _lib
ap = argparse.ArgumentParser(); ap.add_argument("config"); ap.add_argument("--batch", type=int, default=2960)
ap.add_argument("--identical", action="store_true", help="every problem = problem 0: co-resident CTAs stay in phase (instruction-cache experiment)")
a = ap.parse_args()
prob = gen.make_config(a.config, batch=a.batch)
if a.identical:
    rep = lambda v: np.ascontiguousarray(np.repeat(v[:1], a.batch, axis=0))
    prob = sb.BatchProblem(rep(prob.c), rep(prob.A_cm), rep(prob.b), rep(prob.G_cm), rep(prob.h), prob.cones,
                           sing=np.zeros(a.batch, dtype=np.uint8), colmajor=True)
ss = sb.SolverState(prob)
ss.load(prob)
lib = _lib.load()
prm = sb.default_params(path=2)
r = sb.solve_socp_batch(prob, ss, prm, reload=False, want_iterates=False)      # warm
buf32 = (C.c_ulonglong * 32)()
v3 = 16 < prob.n <= 64 and not os.environ.get("SOCP_B200_NO_V3")
fn = lambda b, reset: lib.socp_b200_debug_phase_clocks(ss.handle.ptr, b, reset)
assert fn(buf32, 1) == 0
r = sb.solve_socp_batch(prob, ss, prm, reload=False, want_iterates=False)
assert fn(buf32, 0) == 0
buf = list(buf32)[16:] if v3 else list(buf32)[:16]
names = ["load", "scaling+resid", "hc+head", "syrk", "xtx", "eq", "solve: G cx", "init", "tail", "mid/post", "out",
         "chol_inv", "solve: H n0, K n0", "  (chol: diag factor, warp 0)", "n0 = G'u", "  (chol: wait at barrier 1)"]
if not v3:      # fused_v2 numbers its marks differently: slot 4 = xtx is called "chol" there and slot 11 = chol_inv "xtx"
    pass
tot = sum(buf[i] for i in range(16) if i not in (13, 15))       # 13 and 15 are parts of chol_inv
print(a.config, "batch", a.batch, "solve_ms %.3f" % r.timings["solve_ms"], "mean iters %.2f" % r.iters.mean())
for i, nm in enumerate(names):
    print(f"  {nm:14s} {buf[i]:12d} clk  {100.0*buf[i]/max(tot,1):5.1f}%")
print("  total          %12d clk (CTA 0, all its problems)" % tot)
