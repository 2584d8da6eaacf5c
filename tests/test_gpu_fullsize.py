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


def test_syrk_tma_feed_bit_identical_to_cp_async_feed():
    """The Gram product H = Gt'Gt of the tiled path is fed by TMA (tensor map + mbarrier ring, csrc/syrk_tma.cu); with
    SOCP_B200_NO_TMA it runs the cp.async kernel of linalg.cuh.  Same tiles, same order of the k-blocks: the whole
    solve must not change by a bit.  n = 500 and k = 1000 exercise the zero-filled out-of-bounds columns (500 = 3 x 128
    + 116) and rows (1008 = 50 x 20 + 8) of the TMA boxes."""
    prob = gen.make_config("C4", batch=40)
    a = sb.solve_socp_batch(prob, sb.SolverState(prob))
    os.environ["SOCP_B200_NO_TMA"] = "1"
    try:
        b = sb.solve_socp_batch(prob, sb.SolverState(prob))
    finally:
        del os.environ["SOCP_B200_NO_TMA"]
    assert a.timings["path_used"] == sb.PATH_TILED and (a.status == sb.STATUS_CONVERGED).all()
    for f in ("x", "y", "z", "s", "status", "iters", "pobj", "dobj"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f


GOLD_LARGE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "large_golden.npz")


def test_c4_baseline_batch_vs_golden():
    """BASELINE.json's C4 batch (1000 problems, n=500, k=1000) in one go; 32 of them (every 31st) are pinned by the
    numpy oracle's committed answers (tests/golden/make_golden_large.py): status identical, iterations +-1, objectives
    <= 1e-8 relative where the iteration counts agree; every problem through the size-independent properties."""
    g = np.load(GOLD_LARGE)
    prob = gen.make_config("C4")
    assert prob.B == 1000
    res = sb.solve_socp_batch(prob, sb.SolverState(prob))
    assert res.timings["path_used"] == sb.PATH_TILED
    assert (res.status == sb.STATUS_CONVERGED).all()
    check_properties(prob, res)
    idx = g["c4_index"]
    assert len(idx) >= 32
    assert np.array_equal(res.status[idx], g["c4_status"])
    assert np.all(np.abs(res.iters[idx].astype(int) - g["c4_iters"].astype(int)) <= 1)
    same = res.iters[idx] == g["c4_iters"]
    assert same.sum() >= 24
    rel = lambda a, b: np.abs(a - b) / np.maximum(1.0, np.abs(b))
    d = np.maximum(rel(res.pobj[idx][same], g["c4_pobj"][same]), rel(res.dobj[idx][same], g["c4_dobj"][same]))
    # measured oracle-vs-oracle spread on these 32 problems (C oracle against the numpy oracle's fixture, both on the
    # CPU): 31 within 1e-8, problem 868 (7 iterations) at 1.35e-8 -- the reference algorithm amplifies last-bit
    # differences in its final iterations.  Bar: >= 90 % within 1e-8, all within 5e-8.
    assert d.max() <= 5e-8 and (d <= 1e-8).mean() >= 0.9, (d.max(), (d <= 1e-8).mean())


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
    # the numpy oracle's committed answers for this very problem (block-structured scaling + LAPACK factor and
    # solves, tests/golden/make_golden_large.py): pins src/densesolver.jl:41-90 + the driver at n = 4096
    g = np.load(GOLD_LARGE)
    assert res.status[0] == int(g["c5_status"])
    assert abs(int(res.iters[0]) - int(g["c5_iters"])) <= 1
    if int(res.iters[0]) == int(g["c5_iters"]):
        assert abs(res.pobj[0] - float(g["c5_pobj"])) <= 1e-8 * max(1.0, abs(float(g["c5_pobj"])))
        assert abs(res.dobj[0] - float(g["c5_dobj"])) <= 1e-8 * max(1.0, abs(float(g["c5_dobj"])))
        assert np.max(np.abs(res.x[0] - g["c5_x"])) <= 1e-6 * max(1.0, np.max(np.abs(g["c5_x"])))


def test_c5_step_level_vs_golden():
    """One compute_scaling + setup_iter + solve_kkt at n = 4096, k = 8192 from seeded inputs against the numpy
    oracle's committed outputs (<= 1e-10 relative): src/scalings.jl:32-99, src/densesolver.jl:41-90 at full size."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_large", os.path.join(os.path.dirname(GOLD_LARGE), "make_golden_large.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    g = np.load(GOLD_LARGE)
    prob = gen.make_config("C5")
    cones = cones_t(prob)
    s, z, dx, dz, ds = mg.step_inputs(cones, prob.n, prob.k, 77)
    solver = sb.B200Solver(prob)
    ss = sb.SolverState(prob, solver)
    ss.load(prob)
    sc = sb.compute_scaling(prob.cones, ss.scaling, s, z)
    assert not sc.fail.any()
    rel = lambda a, b: float(np.max(np.abs(a - b)) / max(1.0, np.max(np.abs(b))))
    assert rel(sc.l[0], g["c5_step_lambda"]) < 1e-13
    fail = sb.setup_iter(solver, prob, None, sc)
    assert not fail.any()
    cx, cy, cz, cs = np.zeros((1, prob.n)), np.zeros((1, 0)), np.zeros((1, prob.k)), np.zeros((1, prob.k))
    sb.solve_kkt(solver, prob, None, sc, dx, np.zeros(0), dz, ds, cx, cy, cz, cs)
    assert rel(cx[0], g["c5_step_cx"]) < 1e-10, rel(cx[0], g["c5_step_cx"])
    assert rel(cz[0], g["c5_step_cz"]) < 1e-10, rel(cz[0], g["c5_step_cz"])
    assert rel(cs[0], g["c5_step_cs"]) < 1e-10, rel(cs[0], g["c5_step_cs"])
