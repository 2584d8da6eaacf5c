"""BASELINE.json's configurations at full size on a B200: every problem is checked
through size-independent properties of the domain (the reference's own stop test
recomputed on the host from the returned iterate, strict cone membership of s and
z, weak duality), and a sample is compared with the C oracle (status identical,
iterations within +-1, objectives within 1e-8 relative)."""
import os

import numpy as np
import pytest

import socp_b200 as sb
from socp_b200 import generators as gen
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu


def cones_t(prob):
    return tuple((c.kind, c.offs, c.dim) for c in prob.cones)


def kkt_properties(prob, res, idx=None):
    """Recomputes ||A'y+G'z+c|| + ||Ax-b|| + z's (reference src/solver.jl:122) and the cone
    margins of s and z for the problems in idx (default: all)."""
    B = prob.B
    idx = np.arange(B) if idx is None else np.asarray(idx)
    G = prob.G_cm if prob.shared_G else prob.G_cm[idx]          # (.., n, k) = column-major k x n
    x, z, s = res.x[idx], res.z[idx], res.s[idx]
    if prob.shared_G:
        Gtz = z @ G.T                                            # (b,k)@(k,n)
        Gx = x @ G
    else:
        Gtz = np.einsum("bnk,bk->bn", G, z)
        Gx = np.einsum("bnk,bn->bk", G, x)
    rx = Gtz + prob.c[idx]
    ry = np.zeros((len(idx), 0))
    if prob.p:
        A = prob.A_cm if prob.shared_A else prob.A_cm[idx]
        if prob.shared_A:
            rx = rx + res.y[idx] @ A.T
            ry = x @ A - prob.b[idx]
        else:
            rx = rx + np.einsum("bnp,bp->bn", A, res.y[idx])
            ry = np.einsum("bnp,bn->bp", A, x) - prob.b[idx]
    rz = Gx + s - prob.h[idx]
    gap = np.einsum("bk,bk->b", z, s)
    stop = np.linalg.norm(rx, axis=1) + np.linalg.norm(ry, axis=1) + gap
    margin_s = np.full(len(idx), np.inf)
    margin_z = np.full(len(idx), np.inf)
    for c in prob.cones:
        sl = slice(c.offs, c.offs + c.dim)
        for v, m in ((s, margin_s), (z, margin_z)):
            if c.kind == 0:
                np.minimum(m, v[:, sl].min(axis=1), out=m)
            else:
                np.minimum(m, v[:, c.offs] - np.linalg.norm(v[:, c.offs + 1:c.offs + c.dim], axis=1), out=m)
    return dict(stop=stop, rz=np.abs(rz).max(axis=1), gap=gap, margin_s=margin_s, margin_z=margin_z)


def check_against_oracle(prob, res, sample, strict=True):
    """strict: every sampled objective within 1e-8 relative.  not strict (C3): the reference algorithm
    amplifies 1-ulp differences in late iterations on this shape -- the numpy and C oracles differ from
    EACH OTHER by up to 1.45e-8 on the same sample (3 of 513 above 1e-8, measured; DESIGN.md "parity") --
    so the bar is >= 99% within 1e-8 and all within 1e-7."""
    sample = np.asarray(sample)
    o = co.solve_batch(prob.c[sample], prob.A_cm[sample] if prob.p else np.zeros((len(sample), prob.n, 0)),
                       prob.b[sample], prob.G_cm[sample], prob.h[sample], cones_t(prob),
                       sing=np.zeros(len(sample), np.uint8), nthreads=os.cpu_count())
    assert np.array_equal(res.status[sample], o["status"])
    assert np.all(np.abs(res.iters[sample].astype(int) - o["iters"].astype(int)) <= 1)
    same = res.iters[sample] == o["iters"]
    rel = lambda a, b: np.abs(a - b) / np.maximum(1.0, np.abs(b))
    d = np.maximum(rel(res.pobj[sample][same], o["pobj"][same]), rel(res.dobj[sample][same], o["dobj"][same]))
    if strict:
        assert d.max() <= 1e-8, d.max()
    else:
        assert d.max() <= 1e-7 and (d <= 1e-8).mean() >= 0.99, (d.max(), (d <= 1e-8).mean())
    return int(same.sum())


def check_properties(prob, res, idx=None, tol=1e-5):
    pr = kkt_properties(prob, res, idx)
    st = res.status if idx is None else res.status[np.asarray(idx)]
    conv = st == sb.STATUS_CONVERGED
    # converged <=> the reference's stop test holds on the returned iterate (recomputed in numpy)
    assert np.all(pr["stop"][conv] < tol * (1 + 1e-6) + 1e-12), pr["stop"][conv].max()
    assert np.all(pr["margin_s"][conv] > 0) and np.all(pr["margin_z"][conv] > 0)
    assert np.all(pr["gap"][conv] >= 0)
    # weak duality at the returned point: pobj - dobj = z's + residual terms, small at convergence
    po = res.pobj if idx is None else res.pobj[np.asarray(idx)]
    do = res.dobj if idx is None else res.dobj[np.asarray(idx)]
    assert np.all(np.abs(po[conv] - do[conv]) < 1e-3 * np.maximum(1.0, np.abs(po[conv])))
    return pr


def test_c2_full_batch():
    prob = gen.make_config("C2")                      # 10k portfolio SOCPs, n=50, POC 50 + SOC 51
    ss = sb.SolverState(prob)
    res = sb.solve_socp_batch(prob, ss)
    assert res.timings["path_used"] == sb.PATH_FUSED
    assert (res.status == sb.STATUS_CONVERGED).all()
    assert 5 <= res.iters.min() and res.iters.max() <= 30
    check_properties(prob, res)
    assert check_against_oracle(prob, res, np.arange(0, prob.B, prob.B // 96)) > 80
    # the tiled path must agree with the fused one on the same problems
    sub = gen.make_config("C2", batch=256)
    rt = sb.solve_socp_batch(sub, sb.SolverState(sub), sb.default_params(path=sb.PATH_TILED))
    assert np.array_equal(rt.status, res.status[:256])
    assert np.all(np.abs(rt.iters - res.iters[:256]) <= 1)
    same = rt.iters == res.iters[:256]
    assert np.max(np.abs(rt.pobj[same] - res.pobj[:256][same])) < 1e-8


def test_c3_full_batch():
    prob = gen.make_config("C3")                      # 100k SOCPs, n=12, 10 x SOC(4)
    ss = sb.SolverState(prob)
    res = sb.solve_socp_batch(prob, ss)
    assert res.timings["path_used"] == sb.PATH_FUSED
    assert (res.status == sb.STATUS_CONVERGED).mean() > 0.999
    check_properties(prob, res)
    check_against_oracle(prob, res, np.arange(0, prob.B, prob.B // 512), strict=False)


def test_c4_batch():
    prob = gen.make_config("C4", batch=48)            # n=500, k=1000, 20 x SOC(50): tiled path, DMMA SYRK/Cholesky
    ss = sb.SolverState(prob)
    res = sb.solve_socp_batch(prob, ss)
    assert res.timings["path_used"] == sb.PATH_TILED
    assert (res.status == sb.STATUS_CONVERGED).all()
    check_properties(prob, res)
    check_against_oracle(prob, res, np.arange(0, 48, 6))


def test_c4_many_waves_no_race():
    # more CTAs than one wave per kernel: row-block CTAs of one problem run at different times
    prob = gen.make_config("C4", batch=400)
    res = sb.solve_socp_batch(prob, sb.SolverState(prob), want_iterates=False)
    assert (res.status == sb.STATUS_CONVERGED).all()
    assert res.iters.max() <= 9


def test_c5_single_large():
    prob = gen.make_config("C5")                      # n=4096, k=8192, 64 x SOC(128), one problem
    ss = sb.SolverState(prob)
    res = sb.solve_socp_batch(prob, ss)
    assert res.status[0] == sb.STATUS_CONVERGED and res.iters[0] <= 12
    pr = check_properties(prob, res)
    assert pr["rz"][0] < 1e-6
    # determinism / reuse of the handle
    res2 = sb.solve_socp_batch(prob, ss)
    assert np.array_equal(res.x, res2.x)
