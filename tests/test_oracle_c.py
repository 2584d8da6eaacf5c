"""The C oracle (timed CPU baseline) must agree with the numpy oracle, which is
the one pinned to the reference's golden vectors.  CPU only."""
import numpy as np
import pytest

from oracle import socp_oracle as so
from oracle import c_oracle as co
from socp_b200 import generators as gen
import refcases as rc


def _cones(prob):
    return tuple((c.kind, c.offs, c.dim) for c in prob.cones)


def test_c_kkt_golden():
    g = rc.KKT
    r = co.kkt_step(g["A"], g["G"], g["cones"], False, g["s"], g["z"], g["dx"], g["dy"], g["dz"], g["ds"])
    assert r["rc"] == 0
    assert np.linalg.norm(r["cx"] - g["cxr"]) < 1e-10        # test/runtests.jl:112,124
    assert np.linalg.norm(r["cz"] - g["czr"]) < 1e-10        # :114,126
    assert np.linalg.norm(r["cs"] - g["csr"]) < 1e-10        # :115,127


@pytest.mark.parametrize("name,iters", [("socp1", 5), ("socp2", 12), ("socp3", 10), ("control", 4)])
def test_c_reference_instances(name, iters):
    d = rc.ALL_C1[name]()
    n = d["c"].shape[0]
    r = co.solve_batch(d["c"][None], d["A"].T[None] if d["A"].shape[0] else np.zeros((1, n, 0)), d["b"][None],
                       d["G"].T[None], d["h"][None], d["cones"])
    assert r["status"][0] == so.STATUS_CONVERGED and r["iters"][0] == iters
    if d["xstar"] is not None:
        assert np.linalg.norm(r["x"][0] - d["xstar"]) < 1e-3
    else:
        assert abs(r["pobj"][0] - 0.2901440) < 1e-6


@pytest.mark.parametrize("cfg,B", [("C2", 6), ("C3", 16)])
def test_c_vs_numpy_on_generated(cfg, B):
    prob = gen.make_config(cfg, batch=B)
    r = co.solve_batch(prob.c, prob.A_cm, prob.b, prob.G_cm, prob.h, _cones(prob), sing=prob.sing, nthreads=2)
    for q in range(B):
        pr = so.Problem.create(prob.c[q], prob.A_dense(q), prob.b[q], prob.G_dense(q), prob.h[q], _cones(prob), sing=False)
        ref = so.solve_socp(pr, init="reduced")
        assert r["status"][q] == ref.status and r["iters"][q] == ref.iters
        assert abs(r["pobj"][q] - ref.pobj) <= 1e-8 * max(1.0, abs(ref.pobj))
        assert abs(r["dobj"][q] - ref.dobj) <= 1e-8 * max(1.0, abs(ref.dobj))
        assert np.max(np.abs(r["x"][q] - ref.state.x)) < 1e-4   # late-iteration amplification, SURVEY.md 7.3
