"""Pins the numpy oracle against every golden vector / known answer the
reference's own test file holds for the hot path (/root/reference/test/runtests.jl).
CPU only."""
import math

import numpy as np
import pytest

from oracle import socp_oracle as so
import refcases as rc


# ---- "Vector operations", test/runtests.jl:10-28 ---------------------------------
def test_vprod_exact():
    assert np.array_equal(so.vprod(rc.VEC_CONES, rc.VEC_TV1, rc.VEC_TV1), [1, 1, 1, 14, 4, 6])   # :15
    assert np.array_equal(so.vprod(rc.VEC_CONES, rc.VEC_TV1, rc.VEC_TV2), [1, 1, 1, 29, 7, 9])   # :16


def test_iprod_roundtrip():
    t = so.iprod(rc.VEC_CONES, rc.VEC_TV1, rc.VEC_TV2)
    assert np.linalg.norm(so.vprod(rc.VEC_CONES, rc.VEC_TV1, t) - rc.VEC_TV2) < 1e-4             # :17
    # the O(d) form is the same map
    lam = np.array([1, 2, 3, 7.0, 2, 3])
    assert np.allclose(so.iprod(rc.VEC_CONES, lam, rc.VEC_TV2), so.iprod_fast(rc.VEC_CONES, lam, rc.VEC_TV2),
                       rtol=1e-14, atol=1e-14)


def test_identity_element():
    tid = so.make_e(rc.VEC_CONES)
    assert np.array_equal(so.vprod(rc.VEC_CONES, tid, rc.VEC_TV1), rc.VEC_TV1)                    # :18
    assert np.array_equal(so.vprod(rc.VEC_CONES, tid, rc.VEC_TV2), rc.VEC_TV2)                    # :19


def test_cgt():
    z3 = np.zeros(3)
    assert so.cgt(((so.POC, 0, 3),), np.array([1.0, 2, 3]), z3)                                    # :20
    assert not so.cgt(((so.POC, 0, 3),), np.array([-1.0, 2, 3]), z3)                               # :21
    assert so.cgt(((so.SOC, 0, 3),), np.array([3.0, 2, 2]), z3)                                    # :22
    assert not so.cgt(((so.SOC, 0, 3),), np.array([2.0, 2, 2]), z3)                                # :23


def test_max_step_exact():
    x = np.array([1.0, 2, 3])
    assert so.max_step(((so.POC, 0, 3),), x) == -1                                                 # :25
    assert so.max_step(((so.SOC, 0, 3),), x) == math.sqrt(2 ** 2 + 3 ** 2) - 1.0                   # :26
    assert so.max_step(((so.POC, 0, 3), (so.SOC, 0, 3)), x) == math.sqrt(2 ** 2 + 3 ** 2) - 1.0    # :27


def test_deg():
    assert so.deg(rc.VEC_CONES) == 4
    assert so.deg(((so.SOC, 0, 3), (so.SOC, 3, 4))) == 2


# ---- "Nesterov-Todd Scalings", test/runtests.jl:30-48 -----------------------------
def test_nt_scaling_identities():
    sc = so.Scaling.create(rc.NT_CONES)
    so.compute_scaling(rc.NT_CONES, sc, rc.NT_S, rc.NT_Z)
    sca, isca, pt = sc.W, sc.iW, sc.l
    assert abs(np.sum(isca @ sca - np.eye(6))) < 1e-3                                              # :39
    assert np.linalg.norm(isca.T @ rc.NT_S - sca @ rc.NT_Z) < 1e-3                                 # :40
    assert np.linalg.norm(isca.T @ rc.NT_S - pt) < 1e-3                                            # :41
    op = so.scale(rc.NT_CONES, sc, rc.NT_Z)
    assert np.linalg.norm(sca @ rc.NT_Z - op) < 1e-3                                               # :44
    op2 = so.iscale(rc.NT_CONES, sc, op)
    assert np.linalg.norm(rc.NT_Z - op2) < 1e-3                                                    # :47
    # much tighter than the reference's own 1e-3: these are identities
    assert np.linalg.norm(isca @ sca - np.eye(6)) < 1e-12
    assert np.linalg.norm(isca.T @ rc.NT_S - pt) < 1e-12
    assert np.linalg.norm(sca @ rc.NT_Z - op) < 1e-12
    # closed form W^-2 = eta^-2 (2 v v' - J), v = J wbar   
    wb = sc.wbs[3:6]
    v = wb * np.array([1, -1, -1.0])
    J = np.diag([1, -1, -1.0])
    assert np.allclose(sc.iWiW[3:, 3:], (2 * np.outer(v, v) - J) / sc.mu[1] ** 2, rtol=1e-13, atol=1e-15)


# ---- "Squared NT Scalings", test/runtests.jl:50-93 --------------------------------
@pytest.mark.parametrize("cones,s,z", [
    (rc.NT_CONES, rc.NT_S, rc.NT_Z),        # :51-60
    (rc.SQ_CONES, rc.SQ_U1, rc.SQ_V1),      # :65-73
    (rc.SQ_CONES, rc.SQ_U2, rc.SQ_V2),      # :82-90
])
def test_sqr_scaling_agrees(cones, s, z):
    sc = so.Scaling.create(cones)
    so.compute_scaling(cones, sc, s, z)
    s2 = so.SqrScaling.create(cones)
    so.compute_sqr_scaling(cones, s2, s, z)
    assert np.linalg.norm(sc.l - s2.l) < 1e-4                                                      # :58,:71
    assert np.linalg.norm(sc.wbs - s2.wbs) < 1e-4                                                  # :59,:72
    assert np.linalg.norm(sc.mu - s2.mu) < 1e-4                                                    # :60,:73
    fs = so.compute_full_scaling(cones, s2)
    assert np.linalg.norm(fs - sc.iWiW) < 1e-2                                                     # :80,:89
    assert np.linalg.norm(fs - sc.iWiW) < 1e-12
    assert np.linalg.norm(np.diag(s2.iW) @ np.diag(s2.iW) - np.diag(s2.iWiW)) < 1e-2               # :90


def test_sqr_scaling_reduced_matrix():
    # :74-77: factor of G'(D + uu' - vv')G == factor of G' iWiW G; compared through the inverses
    cones, G = rc.SQ_CONES, rc.SQ_G
    sc = so.Scaling.create(cones)
    so.compute_scaling(cones, sc, rc.SQ_U1, rc.SQ_V1)
    s2 = so.SqrScaling.create(cones)
    so.compute_sqr_scaling(cones, s2, rc.SQ_U1, rc.SQ_V1)
    H1 = G.T @ sc.iWiW @ G
    H2 = G.T @ so.compute_full_scaling(cones, s2) @ G
    assert np.linalg.norm(np.linalg.inv(H2) - np.linalg.inv(H1)) < 1e-2
    assert np.linalg.norm(H2 - H1) < 1e-12


# ---- "KKT reference solution", test/runtests.jl:95-128 ----------------------------
def test_kkt_golden():
    g = rc.KKT
    pr = so.Problem.create(g["c"], g["A"], g["b"], g["G"], g["h"], g["cones"])
    assert pr.sing is False
    sc = so.Scaling.create(pr.cones)
    so.compute_scaling(pr.cones, sc, g["s"], g["z"])
    ds_ = so.DenseSolver(pr)
    ds_.setup_iter(pr, sc)
    cx, cy, cz, cs = ds_.solve_kkt(pr, sc, g["dx"], g["dy"], g["dz"], g["ds"])
    # reference tolerance is 1e-3 (:124-127); the vectors are given to 16 digits
    assert np.linalg.norm(cx - g["cxr"]) < 1e-3
    assert np.linalg.norm(cz - g["czr"]) < 1e-3
    assert np.linalg.norm(cs - g["csr"]) < 1e-3
    assert cy.shape == (0,)
    assert np.linalg.norm(cx - g["cxr"]) < 1e-10
    assert np.linalg.norm(cz - g["czr"]) < 1e-10
    assert np.linalg.norm(cs - g["csr"]) < 1e-10


# ---- end-to-end, test/runtests.jl:130-244 ------------------------------------------
EXPECT = {   # name -> (iterations, sing)
    "socp1": (5, False), "socp2": (12, False), "socp3": (10, False), "control": (4, True),
}


@pytest.mark.parametrize("name", ["socp1", "socp2", "socp3", "control"])
@pytest.mark.parametrize("init", ["full", "reduced"])
def test_solve_socp_reference_instances(name, init):
    d = rc.ALL_C1[name]()
    pr = so.Problem.create(d["c"], d["A"], d["b"], d["G"], d["h"], d["cones"])
    assert pr.sing == EXPECT[name][1]
    res = so.solve_socp(pr, init=init)
    assert res.status == so.STATUS_CONVERGED
    assert res.iters == EXPECT[name][0]
    if d["xstar"] is not None:
        assert np.linalg.norm(res.state.x - d["xstar"]) < 1e-3                                     # :142,:166,:187
    else:
        assert abs(res.pobj - 0.2901440) < 1e-6      # unpinned by the reference; SURVEY.md section 4
    # primal/dual objectives agree at the optimum
    assert abs(res.pobj - res.dobj) < 1e-4
    # cone membership of the final iterate
    zero = np.zeros(pr.k)
    assert so.cgt(pr.cones, res.state.s, zero) and so.cgt(pr.cones, res.state.z, zero)


@pytest.mark.parametrize("name", ["socp1", "socp2", "socp3", "control"])
def test_initial_point_reduced_equals_full(name):
    d = rc.ALL_C1[name]()
    pr = so.Problem.create(d["c"], d["A"], d["b"], d["G"], d["h"], d["cones"])
    xf, yf, zf = so.initial_point_full(pr)
    xr, yr, zr = so.initial_point_reduced(pr)
    assert np.allclose(xf, xr, rtol=1e-9, atol=1e-9)
    assert np.allclose(yf, yr, rtol=1e-9, atol=1e-9)
    assert np.allclose(zf, zr, rtol=1e-9, atol=1e-9)


def test_status_maxiter_and_numerical():
    d = rc.socp2()
    pr = so.Problem.create(d["c"], d["A"], d["b"], d["G"], d["h"], d["cones"])
    res = so.solve_socp(pr, params=so.Params(max_iter=3))
    assert res.status == so.STATUS_MAXITER and res.iters == 3
    # an unbounded problem (drop the constraints that bound it): the reference would throw
    pr2 = so.Problem.create(np.array([-1.0, 0.0]), np.zeros((0, 2)), np.zeros(0),
                            np.array([[0.0, -1.0], [0.0, 0.0], [-1.0, 0.0]]) * 0 + np.array([[0, 0], [0, -1.0], [0, 0]]),
                            np.zeros(3), ((so.SOC, 0, 3),))
    res2 = so.solve_socp(pr2, init="reduced")
    assert res2.status in (so.STATUS_NUMERICAL, so.STATUS_MAXITER)
