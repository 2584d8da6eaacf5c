"""Parity of the CUDA path (through the C ABI) against the oracle, on a B200.
Step level: <= 1e-10 relative from identical inputs.  End to end: status
identical, iterations within +-1, objectives within 1e-8 relative."""
import numpy as np
import pytest

import socp_b200 as sb
from socp_b200 import generators as gen
from oracle import socp_oracle as so
import refcases as rc

pytestmark = pytest.mark.gpu


def ocones(cones):
    return tuple((c.kind, c.offs, c.dim) for c in sb.api._as_cones(cones))


def bcones(cones):
    return tuple(sb.Cone(*c) for c in cones)


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.size == 0 and b.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b)), 1.0))


LAYOUTS = {
    "ref_vec": ((0, 0, 3), (1, 3, 3)),
    "c2": ((0, 0, 50), (1, 50, 51)),
    "c3": tuple((1, 4 * j, 4) for j in range(10)),
    "mixed": ((0, 0, 5), (0, 5, 140), (1, 145, 2), (1, 147, 33), (1, 180, 70)),
    "c4small": tuple((1, 50 * j, 50) for j in range(3)),
}


def interior(cones, rng, B):
    k = so.total_dim(cones)
    v = np.empty((B, k))
    for kind, offs, dim in cones:
        if kind == 0:
            v[:, offs:offs + dim] = rng.uniform(0.5, 2.0, (B, dim))
        else:
            tail = rng.standard_normal((B, dim - 1))
            v[:, offs + 1:offs + dim] = tail
            v[:, offs] = np.linalg.norm(tail, axis=1) + rng.uniform(0.5, 1.5, B)
    return v


# ---------------------------------------------------------------- golden vectors of the reference
def test_vector_ops_golden():
    cones = bcones(rc.VEC_CONES)
    assert np.array_equal(sb.vprod(cones, rc.VEC_TV1, rc.VEC_TV1), [1, 1, 1, 14, 4, 6])      # runtests.jl:15
    assert np.array_equal(sb.vprod(cones, rc.VEC_TV1, rc.VEC_TV2), [1, 1, 1, 29, 7, 9])      # :16
    t = sb.iprod(cones, rc.VEC_TV1, rc.VEC_TV2)
    assert np.linalg.norm(sb.vprod(cones, rc.VEC_TV1, t) - rc.VEC_TV2) < 1e-4                # :17
    tid = sb.make_e(cones)
    assert np.array_equal(sb.vprod(cones, tid, rc.VEC_TV1), rc.VEC_TV1)                       # :18
    assert np.array_equal(sb.vprod(cones, tid, rc.VEC_TV2), rc.VEC_TV2)                       # :19
    x = np.array([1.0, 2, 3])
    assert sb.max_step((sb.POC(0, 3),), x) == -1                                              # :25
    assert abs(sb.max_step((sb.SOC(0, 3),), x) - (np.sqrt(13.0) - 1.0)) < 1e-15               # :26


def test_nt_scaling_golden():
    cones = bcones(rc.NT_CONES)
    prob = sb.Problem(np.zeros(1), np.zeros((0, 1)), np.zeros(0), np.zeros((6, 1)), np.zeros(6), cones, sing=False)
    ss = sb.SolverState(prob, sb.B200Solver(prob))
    sc = sb.compute_scaling(cones, ss.scaling, rc.NT_S, rc.NT_Z)
    osc = so.Scaling.create(rc.NT_CONES)
    so.compute_scaling(rc.NT_CONES, osc, rc.NT_S, rc.NT_Z)
    assert relerr(sc.l[0], osc.l) < 1e-14 and relerr(sc.wbs[0], osc.wbs) < 1e-14 and relerr(sc.mu[0], osc.mu) < 1e-14
    op, op2 = np.zeros(6), np.zeros(6)
    sb.scale_(cones, sc, rc.NT_Z, op)
    assert np.linalg.norm(osc.W @ rc.NT_Z - op) < 1e-12                                       # :44
    assert np.linalg.norm(osc.iW.T @ rc.NT_S - sc.l[0]) < 1e-12                               # :41
    sb.iscale_(cones, sc, op, op2)
    assert np.linalg.norm(rc.NT_Z - op2) < 1e-12                                              # :47
    sb.iwiw(cones, sc, rc.NT_Z, op)
    assert np.linalg.norm(osc.iWiW @ rc.NT_Z - op) < 1e-12


def test_kkt_golden():
    g = rc.KKT
    cones = bcones(g["cones"])
    prob = sb.Problem(g["c"], g["A"], g["b"], g["G"], g["h"], cones)
    solver = sb.B200Solver(prob)
    ss = sb.SolverState(prob, solver)
    ss.load(prob)
    assert ss.get_sing()[0] == 0
    sc = sb.compute_scaling(cones, ss.scaling, g["s"], g["z"])
    st = sb.State(g["x"], g["y"], g["z"], g["s"])
    fail = sb.setup_iter(solver, prob, st, sc)
    assert fail[0] == 0
    cx, cy, cz, cs = np.zeros(3), np.zeros(0), np.zeros(4), np.zeros(4)
    sb.solve_kkt(solver, prob, st, sc, g["dx"], g["dy"], g["dz"], g["ds"], cx, cy, cz, cs)
    assert np.linalg.norm(cx - g["cxr"]) < 1e-10                                              # :124 (ref tol 1e-3)
    assert np.linalg.norm(cz - g["czr"]) < 1e-10                                              # :126
    assert np.linalg.norm(cs - g["csr"]) < 1e-10                                              # :127


# ---------------------------------------------------------------- cone ops vs oracle, random batches
@pytest.mark.parametrize("name", list(LAYOUTS))
def test_cone_ops_vs_oracle(name):
    cones = LAYOUTS[name]
    bc = bcones(cones)
    k = so.total_dim(cones)
    rng = np.random.default_rng(7)
    B = 5
    s, z = interior(cones, rng, B), interior(cones, rng, B)
    u, v = rng.standard_normal((B, k)), rng.standard_normal((B, k))
    prob = sb.BatchProblem(np.zeros((B, 1)), np.zeros((B, 0, 1)), np.zeros((B, 0)), np.zeros((B, k, 1)),
                           np.zeros((B, k)), bc, sing=np.zeros(B, np.uint8))
    ss = sb.SolverState(prob)
    sc = sb.compute_scaling(bc, ss.scaling, s, z)
    assert not sc.fail.any()
    got = dict(vprod=sb.vprod(bc, u, v), iprod=sb.iprod(bc, s, v), max_step=sb.max_step(bc, u),
               step=sb.compute_step(bc, sc.l, u, v), e=sb.make_e(bc, B))
    W, iW, i2 = np.zeros((B, k)), np.zeros((B, k)), np.zeros((B, k))
    sb.scale_(bc, sc, u, W)
    sb.iscale_(bc, sc, u, iW)
    sb.iwiw(bc, sc, u, i2)
    for b in range(B):
        osc = so.Scaling.create(cones)
        so.compute_scaling(cones, osc, s[b], z[b])
        assert relerr(sc.l[b], osc.l) < 1e-13
        assert relerr(sc.wbs[b], osc.wbs) < 1e-13
        assert relerr(sc.mu[b], osc.mu) < 1e-13
        assert relerr(got["vprod"][b], so.vprod(cones, u[b], v[b])) < 1e-13
        assert relerr(got["iprod"][b], so.iprod(cones, s[b], v[b])) < 1e-11
        assert abs(got["max_step"][b] - so.max_step(cones, u[b])) < 1e-13
        assert abs(got["step"][b] - so.compute_step(cones, osc.l, u[b], v[b])) < 1e-12
        assert np.array_equal(got["e"][b], so.make_e(cones))
        assert relerr(W[b], so.scale(cones, osc, u[b])) < 1e-13
        assert relerr(iW[b], so.iscale(cones, osc, u[b])) < 1e-13
        assert relerr(i2[b], osc.iWiW @ u[b]) < 1e-12


def test_scaling_failure_flag():
    cones = LAYOUTS["c3"]
    bc = bcones(cones)
    rng = np.random.default_rng(1)
    s, z = interior(cones, rng, 3), interior(cones, rng, 3)
    s[1, 4] = 0.0          # head of cone 1 -> outside the cone: sqrt of a negative in the reference
    prob = sb.BatchProblem(np.zeros((3, 1)), np.zeros((3, 0, 1)), np.zeros((3, 0)), np.zeros((3, 40, 1)),
                           np.zeros((3, 40)), bc, sing=np.zeros(3, np.uint8))
    ss = sb.SolverState(prob)
    sc = sb.compute_scaling(bc, ss.scaling, s, z)
    assert list(sc.fail != 0) == [False, True, False]


# ---------------------------------------------------------------- factor + solve vs oracle
def _kkt_case(n, p, cones, B, seed, sing_mode=False):
    rng = np.random.default_rng(seed)
    k = so.total_dim(cones)
    G = rng.standard_normal((B, k, n)) / np.sqrt(n)
    if sing_mode:
        G[:, :, : n // 3] = 0.0            # rank deficient: G'G singular -> sing = true
    A = rng.standard_normal((B, p, n)) / np.sqrt(n)
    return G, A, rng


@pytest.mark.parametrize("n,p,lay,sing_mode", [
    (3, 0, "ref_vec", False), (50, 1, "c2", False), (12, 0, "c3", False), (70, 5, "mixed", False),
    (130, 0, "c4small", False), (40, 30, "c4small", True), (201, 7, "mixed", False),
])
def test_factor_solve_vs_oracle(n, p, lay, sing_mode):
    cones = LAYOUTS[lay]
    bc = bcones(cones)
    k = so.total_dim(cones)
    B = 3
    G, A, rng = _kkt_case(n, p, cones, B, 11 + n, sing_mode)
    c, b, h = rng.standard_normal((B, n)), rng.standard_normal((B, p)), rng.standard_normal((B, k))
    s, z = interior(cones, rng, B), interior(cones, rng, B)
    dx, dy = rng.standard_normal((B, n)), rng.standard_normal((B, p))
    dz, ds = rng.standard_normal((B, k)), rng.standard_normal((B, k))
    prob = sb.BatchProblem(c, A, b, G, h, bc)
    solver = sb.B200Solver(prob)
    ss = sb.SolverState(prob, solver)
    ss.load(prob)
    sing = ss.get_sing()
    sc = sb.compute_scaling(bc, ss.scaling, s, z)
    fail = sb.setup_iter(solver, prob, None, sc)
    assert not fail.any()
    cx, cy, cz, cs = np.zeros((B, n)), np.zeros((B, p)), np.zeros((B, k)), np.zeros((B, k))
    sb.solve_kkt(solver, prob, None, sc, dx, dy, dz, ds, cx, cy, cz, cs)
    for q in range(B):
        pr = so.Problem.create(c[q], A[q], b[q], G[q], h[q], cones)
        assert bool(sing[q]) == pr.sing == sing_mode
        osc = so.Scaling.create(cones)
        so.compute_scaling(cones, osc, s[q], z[q])
        dsol = so.DenseSolver(pr)
        dsol.setup_iter(pr, osc)
        ox, oy, oz, os_ = dsol.solve_kkt(pr, osc, dx[q], dy[q], dz[q], ds[q], fast_iprod=True)
        tol = 1e-10 if not sing_mode else 1e-8
        assert relerr(cx[q], ox) < tol, ("cx", relerr(cx[q], ox))
        assert relerr(cy[q], oy) < tol, ("cy", relerr(cy[q], oy))
        assert relerr(cz[q], oz) < tol, ("cz", relerr(cz[q], oz))
        assert relerr(cs[q], os_) < tol, ("cs", relerr(cs[q], os_))


# ---------------------------------------------------------------- end to end, the reference's instances (C1)
EXPECT_ITERS = {"socp1": 5, "socp2": 12, "socp3": 10, "control": 4}


@pytest.mark.parametrize("name", ["socp1", "socp2", "socp3", "control"])
@pytest.mark.parametrize("path", [sb.PATH_TILED, sb.PATH_AUTO])
def test_reference_instances(name, path):
    d = rc.ALL_C1[name]()
    cones = bcones(d["cones"])
    prob = sb.Problem(d["c"], d["A"], d["b"], d["G"], d["h"], cones)
    ss = sb.SolverState(prob, sb.B200Solver(prob))
    st = sb.solve_socp(prob, ss, sb.default_params(path=path))
    pr = so.Problem.create(d["c"], d["A"], d["b"], d["G"], d["h"], d["cones"])
    ref = so.solve_socp(pr, init="full")
    assert st.status == ref.status == sb.STATUS_CONVERGED
    assert st.iters == ref.iters == EXPECT_ITERS[name]
    if d["xstar"] is not None:
        assert np.linalg.norm(st.x - d["xstar"]) < 1e-3          # runtests.jl:142,166,187
    # SOCP2/SOCP3 sit at the edge of what the reference algorithm resolves: 1-ulp
    # reformulations of the ORACLE move x by 1e-4 there (DESIGN.md "parity"), so the
    # bar is the reference's own (x to 1e-3, :142) plus objectives to 1e-6; the
    # well-conditioned instances are held to 1e-8.
    otol = 1e-8 if name in ("socp1", "control") else 1e-6
    assert abs(st.pobj - ref.pobj) <= otol * max(1.0, abs(ref.pobj))
    assert abs(st.dobj - ref.dobj) <= otol * max(1.0, abs(ref.dobj))
    assert relerr(st.x, ref.state.x) < 1e-3


def _check_batch(prob, res, sample, fast=True):
    worst = dict(pobj=0.0, dobj=0.0, x=0.0)
    for q in sample:
        pr = so.Problem.create(prob.c[q], prob.A_dense(q), prob.b[q], prob.G_dense(q), prob.h[q], ocones(prob.cones),
                               sing=False if prob.sing is not None and not prob.sing[q] else None)
        ref = so.solve_socp(pr, init="reduced", fast_iprod=fast)
        assert res.status[q] == ref.status, (q, res.status[q], ref.status)
        assert abs(int(res.iters[q]) - ref.iters) <= 1, (q, res.iters[q], ref.iters)
        if ref.status == so.STATUS_CONVERGED and res.iters[q] == ref.iters:
            worst["pobj"] = max(worst["pobj"], abs(res.pobj[q] - ref.pobj) / max(1.0, abs(ref.pobj)))
            worst["dobj"] = max(worst["dobj"], abs(res.dobj[q] - ref.dobj) / max(1.0, abs(ref.dobj)))
            worst["x"] = max(worst["x"], relerr(res.x[q], ref.state.x))
    assert worst["pobj"] <= 1e-8 and worst["dobj"] <= 1e-8, worst
    return worst


@pytest.mark.parametrize("cfg,B,path", [("C2", 48, sb.PATH_TILED), ("C3", 64, sb.PATH_TILED),
                                        ("C2", 48, sb.PATH_AUTO), ("C3", 64, sb.PATH_AUTO)])
def test_batch_vs_oracle(cfg, B, path):
    prob = gen.make_config(cfg, batch=B)
    ss = sb.SolverState(prob)
    res = sb.solve_socp_batch(prob, ss, sb.default_params(path=path))
    assert (res.status == sb.STATUS_CONVERGED).all()
    _check_batch(prob, res, range(0, B, 3))


def test_mid_size_multi_panel():
    # n = 150 > one Cholesky panel; random feasible with equalities
    cones = tuple((1, 40 * j, 40) for j in range(5))
    prob = gen.random_feasible(4, 150, 6, bcones(cones), 0.05)
    ss = sb.SolverState(prob)
    res = sb.solve_socp_batch(prob, ss)
    assert (res.status == sb.STATUS_CONVERGED).all()
    _check_batch(prob, res, range(4))


def test_maxiter_and_reuse():
    prob = gen.make_config("C3", batch=8)
    ss = sb.SolverState(prob)
    r3 = sb.solve_socp_batch(prob, ss, sb.default_params(max_iter=3))
    assert (r3.status == sb.STATUS_MAXITER).all() and (r3.iters == 3).all()
    # the same SolverState is reusable (reference test/runtests.jl:243) and deterministic
    r1 = sb.solve_socp_batch(prob, ss)
    r2 = sb.solve_socp_batch(prob, ss)
    assert np.array_equal(r1.x, r2.x) and np.array_equal(r1.iters, r2.iters)


def test_shared_G_batch():
    base = gen.make_config("C3", batch=1)
    B = 6
    rng = np.random.default_rng(3)
    h = base.h[0] + 0.01 * np.abs(rng.standard_normal((B, base.k))) * (np.arange(base.k) % 4 == 0)
    c = np.repeat(base.c, B, axis=0)
    shared = sb.BatchProblem(c, np.zeros((B, 0, base.n)), np.zeros((B, 0)), base.G_dense(0), h, base.cones,
                             sing=np.zeros(B, np.uint8))
    full = sb.BatchProblem(c, np.zeros((B, 0, base.n)), np.zeros((B, 0)), np.repeat(base.G_dense(0)[None], B, 0), h,
                           base.cones, sing=np.zeros(B, np.uint8))
    for path in (sb.PATH_TILED, sb.PATH_AUTO):
        ra = sb.solve_socp_batch(shared, sb.SolverState(shared), sb.default_params(path=path))
        rb = sb.solve_socp_batch(full, sb.SolverState(full), sb.default_params(path=path))
        assert ra.timings["path_used"] == (sb.PATH_TILED if path == sb.PATH_TILED else sb.PATH_FUSED)
        assert (ra.status == sb.STATUS_CONVERGED).all()
        assert np.array_equal(ra.status, rb.status) and np.array_equal(ra.x, rb.x) and np.array_equal(ra.s, rb.s)


def _csc_of(prob):
    """The reference's storage of one batch (SparseMatrixCSC, src/Socp.jl:25,29): the union pattern of the
    batch's dense G (and A) with per-problem values, built with scipy on the host -- test-side only."""
    import scipy.sparse as sp
    out = []
    for M_cm, shared, rows in ((prob.G_cm, prob.shared_G, prob.k), (prob.A_cm, prob.shared_A, prob.p)):
        if rows == 0:
            out.append(None)
            continue
        M = M_cm[None] if shared else M_cm                      # (B, n, rows): column-major per problem
        pat = sp.csc_matrix((np.abs(M).max(axis=0) != 0).T.astype(np.float64))
        pat.sort_indices()
        cols = np.repeat(np.arange(prob.n), np.diff(pat.indptr))
        vals = M[:, cols, pat.indices]                          # (B, nnz) in CSC order
        out.append((pat, vals[0] if shared else np.ascontiguousarray(vals)))
    return out


@pytest.mark.parametrize("cfg,B,base", [("C2", 300, 1), ("C3", 500, 0), ("P1", 64, 1)])
def test_csc_ingestion_bit_exact(cfg, B, base):
    """socp_b200_set_data_csc (the reference's SparseMatrixCSC fields in, dense operands assembled on the
    device) must give bit-identical results to socp_b200_set_data with the densified matrices."""
    if cfg == "P1":
        dense = gen.random_feasible(B, 20, 3, (sb.POC(0, 6), sb.SOC(6, 5), sb.SOC(11, 9)), 0.1)
    else:
        dense = gen.make_config(cfg, batch=B)
    (Gp, Gv), A_ = _csc_of(dense)
    def mk(pat, vals):
        return sb.CscMatrix(pat.shape, pat.indptr + base, pat.indices + base, vals, index_base=base)
    G = mk(Gp, Gv)
    A = mk(*A_) if A_ is not None else np.zeros((B, 0, dense.n))
    if cfg == "C2":
        assert G.nnz == 50 + 50 * 50 and A.nnz == 50           # -I, F, and the budget row: half of G is structural zero
    sparse = sb.BatchProblem(dense.c, A, dense.b, G, dense.h, dense.cones, sing=dense.sing)
    for path in (sb.PATH_AUTO, sb.PATH_TILED):
        rd = sb.solve_socp_batch(dense, sb.SolverState(dense), sb.default_params(path=path))
        rs = sb.solve_socp_batch(sparse, sb.SolverState(sparse), sb.default_params(path=path))
        for f in ("x", "y", "z", "s", "status", "iters", "pobj", "dobj"):
            assert np.array_equal(getattr(rd, f), getattr(rs, f)), (path, f)
        if cfg != "P1":                                          # the mixed-cone family is fragile under the reference's
            assert (rs.status == sb.STATUS_CONVERGED).all()      # algorithm; equality is the point
        else:
            assert (rs.status == sb.STATUS_CONVERGED).sum() >= B // 3


def test_csc_shared_pattern_and_values():
    """One SparseMatrixCSC shared by the whole batch (SOCP_FLAG_SHARED_G) == the dense shared-G upload;
    `sing` computed on the device from the assembled matrix."""
    base = gen.make_config("C3", batch=1)
    import scipy.sparse as sp
    B = 6
    rng = np.random.default_rng(3)
    Gd = base.G_dense(0) * (rng.random((base.k, base.n)) < 0.4)
    Gd[np.arange(base.n), np.arange(base.n)] += 1.0             # keep full column rank
    e = (np.arange(base.k) % 4 == 0).astype(np.float64)      # identity element of 10 x SOC(4)
    h = e[None] * (1.0 + 0.1 * np.arange(B))[:, None]           # x = 0, s = h strictly inside the cone
    c = np.repeat((-Gd.T @ e)[None], B, axis=0)                 # z = e strictly dual feasible
    none = (np.zeros((B, 0, base.n)), np.zeros((B, 0)))
    dense = sb.BatchProblem(c, *none, Gd, h, base.cones)
    sparse = sb.BatchProblem(c, *none, sb.CscMatrix.from_scipy(sp.csc_matrix(Gd)), h, base.cones)
    assert sparse.shared_G and sparse.G_csc.nnz < base.k * base.n // 2
    for path in (sb.PATH_TILED, sb.PATH_AUTO):
        ssd, sss = sb.SolverState(dense), sb.SolverState(sparse)
        rd = sb.solve_socp_batch(dense, ssd, sb.default_params(path=path))
        rs = sb.solve_socp_batch(sparse, sss, sb.default_params(path=path))
        assert rs.timings["path_used"] == (sb.PATH_TILED if path == sb.PATH_TILED else sb.PATH_FUSED)
        assert np.array_equal(ssd.get_sing(), sss.get_sing())
        assert (rs.status == sb.STATUS_CONVERGED).all()
        assert np.array_equal(rd.status, rs.status) and np.array_equal(rd.iters, rs.iters)
        assert np.max(np.abs(rd.x - rs.x)) <= 1e-9 * max(1.0, np.max(np.abs(rd.x)))


def test_csc_rejects_broken_patterns():
    """The SparseMatrixCSC invariants are the usage contract: violations are SOCP_ERR_LAYOUT / _SIZE, not UB."""
    prob = gen.make_config("C3", batch=2)
    (Gp, Gv), _ = _csc_of(prob)
    none = (np.zeros((2, 0, prob.n)), np.zeros((2, 0)))
    def attempt(colptr, rowval, vals, base=0):
        G = sb.CscMatrix(Gp.shape, colptr, rowval, vals, index_base=base)
        bad = sb.BatchProblem(prob.c, *none, G, prob.h, prob.cones)
        sb.SolverState(bad).load(bad)
    attempt(Gp.indptr, Gp.indices, Gv)                           # the intact pattern loads
    rv = Gp.indices.copy(); rv[[0, 1]] = rv[[1, 0]]              # unsorted rows inside a column
    with pytest.raises(sb.SocpError):
        attempt(Gp.indptr, rv, Gv)
    rv = Gp.indices.copy(); rv[1] = rv[0]                        # duplicate entry
    with pytest.raises(sb.SocpError):
        attempt(Gp.indptr, rv, Gv)
    rv = Gp.indices.copy(); rv[-1] = prob.k                      # row out of range
    with pytest.raises(sb.SocpError):
        attempt(Gp.indptr, rv, Gv)
    cp = Gp.indptr.copy(); cp[-1] -= 1                           # colptr does not span nnz
    with pytest.raises(sb.SocpError):
        attempt(cp, Gp.indices, Gv)
    with pytest.raises(sb.SocpError):                            # 1-based arrays declared 0-based
        attempt(Gp.indptr + 1, Gp.indices + 1, Gv, base=0)


def test_solve_host_pipelined_matches_two_step():
    """socp_b200_solve_host (chunked upload / solve / download on three streams) must return exactly what
    set_data + solve return: same kernel, same data, only the schedule differs."""
    prob = gen.make_config("C2", batch=5000)          # 2 chunks on a 148-SM part (8 waves of 296 CTAs each)
    ss = sb.SolverState(prob)
    one = sb.solve_socp_batch(prob, ss)                # reload=True -> socp_b200_solve_host
    assert one.timings["path_used"] == sb.PATH_FUSED and one.timings["kernel_launches"] >= 2
    ss2 = sb.SolverState(prob)
    ss2.load(prob)
    two = sb.solve_socp_batch(prob, ss2, reload=False) # socp_b200_set_data + socp_b200_solve
    for f in ("x", "y", "z", "s", "status", "iters", "pobj", "dobj"):
        assert np.array_equal(getattr(one, f), getattr(two, f)), f
    assert (one.status == sb.STATUS_CONVERGED).all()


@pytest.mark.parametrize("sing_known", [True, False])
def test_solve_host_csc_pipelined_matches_dense(sing_known):
    """socp_b200_solve_host_csc -- Problem(...) + solve_socp from the reference's own SparseMatrixCSC storage in one
    pipelined call (values of every chunk uploaded, scattered into the dense operands on the device, solved, downloaded)
    -- must return exactly what the dense one-shot call and the two-step CSC path return."""
    dense = gen.make_config("C2", batch=5000)          # several chunks on a 148-SM part
    (Gp, Gv), (Ap, Av) = _csc_of(dense)
    G = sb.CscMatrix(Gp.shape, Gp.indptr + 1, Gp.indices + 1, Gv, index_base=1)      # 1-based, as Julia hands them over
    A = sb.CscMatrix(Ap.shape, Ap.indptr + 1, Ap.indices + 1, Av, index_base=1)
    assert G.nnz == 2550
    sing = dense.sing if sing_known else None
    sparse = sb.BatchProblem(dense.c, A, dense.b, G, dense.h, dense.cones, sing=sing)
    dense2 = sb.BatchProblem(dense.c, dense.A_cm, dense.b, dense.G_cm, dense.h, dense.cones, sing=sing, colmajor=True)
    one = sb.solve_socp_batch(sparse, sb.SolverState(sparse))          # socp_b200_solve_host_csc
    assert one.timings["path_used"] == sb.PATH_FUSED and one.timings["kernel_launches"] >= 4    # scatters + solves
    ref = sb.solve_socp_batch(dense2, sb.SolverState(dense2))          # socp_b200_solve_host
    ss2 = sb.SolverState(sparse)
    ss2.load(sparse)                                                   # socp_b200_set_data_csc
    two = sb.solve_socp_batch(sparse, ss2, reload=False)               # socp_b200_solve
    for f in ("x", "y", "z", "s", "status", "iters", "pobj", "dobj"):
        assert np.array_equal(getattr(one, f), getattr(ref, f)), f
        assert np.array_equal(getattr(one, f), getattr(two, f)), f
    assert (one.status == sb.STATUS_CONVERGED).all()


def test_multi_device_handle_matches_single_device():
    """socp_b200_create(devices=[0, 1]) shards the batch contiguously over the devices of one process (no collective);
    results must equal the single-device ones bit for bit (same kernels, same problems).  On a box with one GPU the
    second shard lives on device 0 as well (the same for_each_shard code path: one host worker + stream per shard)."""
    import torch
    d1 = 1 if torch.cuda.device_count() >= 2 else 0
    prob = gen.make_config("C2", batch=301)           # odd split: 151 + 150
    one = sb.solve_socp_batch(prob, sb.SolverState(prob, devices=[0]))
    two = sb.solve_socp_batch(prob, sb.SolverState(prob, devices=[0, d1]))
    for f in ("x", "y", "z", "s", "status", "iters", "pobj", "dobj"):
        assert np.array_equal(getattr(one, f), getattr(two, f)), f
    prob3 = gen.make_config("C3", batch=1000)
    a = sb.solve_socp_batch(prob3, sb.SolverState(prob3, devices=[d1]))
    b = sb.solve_socp_batch(prob3, sb.SolverState(prob3, devices=[0, d1]))
    assert np.array_equal(a.status, b.status) and np.array_equal(a.x, b.x)


def test_multi_shard_handle_on_one_device():
    """The same sharding code path (for_each_shard: one host worker + stream per shard, results gathered by shard
    offset) on a box with ONE GPU: devices=[0, 0, 0] makes three shards of the batch on device 0.  Must equal the
    single-shard results bit for bit."""
    prob = gen.make_config("C2", batch=301)           # odd split: 101 + 101 + 99
    one = sb.solve_socp_batch(prob, sb.SolverState(prob, devices=[0]))
    three = sb.solve_socp_batch(prob, sb.SolverState(prob, devices=[0, 0, 0]))
    for f in ("x", "y", "z", "s", "status", "iters", "pobj", "dobj"):
        assert np.array_equal(getattr(one, f), getattr(three, f)), f
    # tiled path + the device-side `sing` test (sing=None) through the shards as well
    p3 = gen.random_feasible(50, 20, 2, (sb.POC(0, 6), sb.SOC(6, 20)), 0.1, sing_known=False)
    a = sb.solve_socp_batch(p3, sb.SolverState(p3, devices=[0]), sb.default_params(path=sb.PATH_TILED))
    b = sb.solve_socp_batch(p3, sb.SolverState(p3, devices=[0, 0]), sb.default_params(path=sb.PATH_TILED))
    assert np.array_equal(a.status, b.status) and np.array_equal(a.x, b.x)


def test_tiled_path_more_than_65535_problems():
    """gridDim.y is capped at 65535: the batch-wide tiled kernels fold the batch over grid y and z.  70 000 tiny
    problems through the tiled path with `sing` computed on the device must equal the fused path's outcomes."""
    B = 70_000
    base = gen.random_feasible(512, 4, 1, (sb.POC(0, 2), sb.SOC(2, 3)), 0.1, sing_known=False)
    rep = (B + 511) // 512
    tile = lambda a: np.ascontiguousarray(np.concatenate([a] * rep, axis=0)[:B])
    prob = sb.BatchProblem(tile(base.c), tile(base.A_cm), tile(base.b), tile(base.G_cm), tile(base.h), base.cones,
                           sing=None, colmajor=True)
    ss = sb.SolverState(prob)
    ss.load(prob)
    assert not ss.get_sing().any()
    til = sb.solve_socp_batch(prob, ss, sb.default_params(path=sb.PATH_TILED), reload=False)
    fus = sb.solve_socp_batch(prob, ss, sb.default_params(path=sb.PATH_FUSED), reload=False)
    assert til.timings["path_used"] == sb.PATH_TILED and fus.timings["path_used"] == sb.PATH_FUSED
    # the batch is 512 distinct problems repeated: every repetition must give the same answer as the first
    for r in (til, fus):
        assert np.array_equal(r.status.reshape(-1)[:(B // 512) * 512].reshape(-1, 512), np.broadcast_to(r.status[:512], (B // 512, 512)))
        assert np.array_equal(r.pobj[:(B // 512) * 512].reshape(-1, 512), np.broadcast_to(r.pobj[:512], (B // 512, 512)))
    conv = fus.status == sb.STATUS_CONVERGED
    assert conv.mean() > 0.5
    assert np.all(til.status[conv] == sb.STATUS_CONVERGED)
    same = conv & (til.iters == fus.iters)
    assert np.max(np.abs(til.pobj[same] - fus.pobj[same]) / np.maximum(1.0, np.abs(fus.pobj[same]))) <= 1e-5


GENERIC_LAYOUTS = {
    # name: (n, p, cones) -- exercise every generic variant of the fused kernel (1, 4 and 8 warps per problem),
    # equality rows handled by one warp (p <= 8) and by the blocked factorisation (p > 8), LP-only and SOC-only
    # layouts, and degenerate cone dimensions
    "lp_only": (10, 3, (sb.POC(0, 25),)),
    "mixed_p2": (20, 2, (sb.POC(0, 6), sb.SOC(6, 9), sb.SOC(15, 5))),
    "soc_n40": (40, 0, gen.soc_cones(4, 12)),
    "mixed_n60_p9": (60, 9, (sb.POC(0, 10), sb.SOC(10, 30), sb.SOC(40, 30), sb.SOC(70, 30))),
    "tiny_cones": (8, 1, (sb.POC(0, 3), sb.SOC(3, 1), sb.SOC(4, 2), sb.SOC(6, 3), sb.SOC(9, 7))),
    "many_small": (30, 4, gen.soc_cones(20, 3)),
    # the limits of the fused kernel (f2_plan): n = 64, p = 32, a 128-dimensional cone (32 lanes x 4 elements)
    "edge_n64": (64, 0, (sb.POC(0, 16), sb.SOC(16, 64))),
    "edge_p32": (40, 32, (sb.POC(0, 50), sb.SOC(50, 20))),
    "edge_soc128": (24, 0, (sb.SOC(0, 128),)),
}


@pytest.mark.parametrize("name", list(GENERIC_LAYOUTS))
def test_fused_generic_layouts_vs_c_oracle(name):
    from oracle import c_oracle as co
    n, p, cones = GENERIC_LAYOUTS[name]
    B = 24
    prob = gen.random_feasible(B, n, p, cones, 0.1)
    ss = sb.SolverState(prob)
    res = sb.solve_socp_batch(prob, ss)
    assert res.timings["path_used"] == sb.PATH_FUSED
    ref = co.solve_batch(prob.c, prob.A_cm, prob.b, prob.G_cm, prob.h, ocones(prob.cones), sing=prob.sing, nthreads=4)
    # the same problems converge; where the reference algorithm breaks down (it would throw), the iteration at which
    # it does is rounding dependent (the numpy and C oracles differ there too), so only "did not converge" is compared
    conv = ref["status"] == sb.STATUS_CONVERGED
    agree = (res.status == sb.STATUS_CONVERGED) == conv
    if not agree.all():
        # A flip is tolerated only on a problem where the reference algorithm itself is rounding-limited: the numpy
        # oracle (sparse block scaling) and the C oracle (dense formulation) must already disagree there -- on status,
        # on the iteration count, or by more than 1e-8 in the final x (the error growth per iteration near the end of
        # these badly conditioned families is ~1e3, so the stop test of the next iteration is a coin toss).
        from oracle import socp_oracle as so
        assert (~agree).sum() <= 2, (res.status, ref["status"])
        for q in np.flatnonzero(~agree):
            pr = so.Problem.create(prob.c[q], prob.A_dense(q), prob.b[q], prob.G_dense(q), prob.h[q], ocones(prob.cones), sing=False)
            r2 = so.solve_socp(pr, init="reduced", fast_iprod=True)
            xd = np.max(np.abs(r2.state.x - ref["x"][q])) / max(1.0, np.max(np.abs(ref["x"][q])))
            assert r2.status != ref["status"][q] or r2.iters != ref["iters"][q] or xd > 1e-8, (q, xd)
    conv = conv & agree
    assert np.all(np.abs(res.iters[conv].astype(int) - ref["iters"][conv].astype(int)) <= 1)
    same = (res.iters == ref["iters"]) & conv
    assert same.sum() >= B // 2
    # These families stop at the reference's loose absolute test (1e-5) on badly conditioned KKT systems: measured
    # numpy-oracle vs C-oracle spread of the objectives on the same problems: mixed_p2 max 2.2e-7, median 2.8e-8 (29 %
    # within 1e-8); mixed_n60_p9 max 7.6e-7.  The bar is therefore max 1e-5 and median 1e-7; the BASELINE.json shapes
    # are held to 1e-8 throughout (test_batch_vs_oracle, tests/test_gpu_fullsize.py).
    rel = lambda a, b: np.abs(a - b) / np.maximum(1.0, np.abs(b))
    d = np.maximum(rel(res.pobj[same], ref["pobj"][same]), rel(res.dobj[same], ref["dobj"][same]))
    assert d.max() <= 1e-5 and np.median(d) <= 1e-7, (d.max(), np.median(d))
    # The tiled path takes H = (W^-1 G)'(W^-1 G), a Gram matrix, where the fused kernel (and the reference, and both
    # oracles) take G' (W^-2 G): when W is so badly conditioned that the latter loses positive definiteness and
    # cholesky! throws, the Gram form can still factor -- the tiled path converges on a superset.
    til = sb.solve_socp_batch(prob, sb.SolverState(prob), sb.default_params(path=sb.PATH_TILED))
    assert np.all(til.status[conv] == sb.STATUS_CONVERGED)
    both = (til.iters == res.iters) & conv
    dt = rel(til.pobj[both], res.pobj[both])
    assert dt.max() <= 1e-5 and np.median(dt) <= 1e-7


# ------------------------------------------------------------------------------------------------ fused_v3 kernel
STEP_LAYOUTS = {
    "C2": None,
    "mixed_p2": (20, 2, (sb.POC(0, 6), sb.SOC(6, 9), sb.SOC(15, 5))),
    "mixed_n60_p9": (60, 9, (sb.POC(0, 10), sb.SOC(10, 30), sb.SOC(40, 30), sb.SOC(70, 30))),
    "edge_p32": (40, 32, (sb.POC(0, 50), sb.SOC(50, 20))),
    "edge_n64": (64, 0, (sb.POC(0, 16), sb.SOC(16, 64))),
}


@pytest.mark.parametrize("name", list(STEP_LAYOUTS))
def test_fused_step_level_vs_oracle(name):
    """One Mehrotra step taken out of the fused whole-solve kernel (socp_b200_debug_fused_step) against
    DenseSolver.setup_iter / solve_kkt of the numpy oracle (reference src/densesolver.jl:41-90) from the kernel's own
    (s, z) and right-hand sides: H = G'W^-2 G <= 1e-12, cx, cy, cz, cs <= 1e-10 on C2 (1e-9 on the generated families,
    whose systems are worse conditioned).  Pins the diagonal + low-rank assembly of the KKT matrix (SURVEY.md section
    8, row f4) and the fused factor / solve, which the whole-solve tests only see end to end."""
    from oracle import socp_oracle as so
    if STEP_LAYOUTS[name] is None:
        prob = gen.make_config("C2", batch=8)
        tol = 1e-10
    else:
        n, p, cones = STEP_LAYOUTS[name]
        prob = gen.random_feasible(8, n, p, cones, 0.1)
        tol = 1e-9
    ss = sb.SolverState(prob)
    ss.load(prob)
    nrm = lambda a, b: np.max(np.abs(a - b)) / np.max(np.abs(b))
    # later iterations of the generated families are too badly conditioned for a 1e-9 bar (cond(H) grows ~30x per step)
    for q, it, phase in (((0, 0, 1), (3, 2, 2), (7, 4, 1)) if STEP_LAYOUTS[name] is None else ((0, 0, 1), (3, 1, 2), (7, 2, 1))):
        d = ss.debug_fused_step(q, it, phase)
        pr = so.Problem.create(prob.c[q], prob.A_dense(q), prob.b[q], prob.G_dense(q), prob.h[q], ocones(prob.cones), sing=False)
        sc = so.compute_scaling(pr.cones, so.Scaling.create(pr.cones), d["s"], d["z"])
        dsv = so.DenseSolver(pr)
        dsv.setup_iter(pr, sc)
        assert nrm(d["H"], dsv.H) <= 1e-12, (q, it, nrm(d["H"], dsv.H))
        cx, cy, cz, cs = dsv.solve_kkt(pr, sc, d["dx"], d["dy"], d["dz"], d["ds"], fast_iprod=True)
        for nm, v in (("cx", cx), ("cy", cy), ("cz", cz), ("cs", cs)):
            if v.size:
                assert nrm(d[nm], v) <= tol, (q, it, phase, nm, nrm(d[nm], v))


def _sing_mask(n, k, zero_cols):
    m = np.ones((k, n), dtype=bool)
    m[:, list(zero_cols)] = False
    return m


def test_mixed_sing_batch_through_solve_host():
    """`sing` on the production entry (reference src/Socp.jl:49-56): socp_b200_solve_host with sing == NULL on a batch
    where every second problem has a rank-deficient G (two variables appear in no cone row; the equality rows pin them
    down).  The fused kernel's initial factorisation of G'G is the reference's test; the flagged problems are solved
    with A'A added (src/densesolver.jl:44-46,68-80).  Statuses / objectives against the C oracle told which are sing."""
    from oracle import c_oracle as co
    n, p = 20, 3
    cones = (sb.POC(0, 8), sb.SOC(8, 14))
    B = 64
    reg = gen.random_feasible_pattern(B, n, p, cones, np.ones((22, n), dtype=bool), 0.1)
    sng = gen.random_feasible_pattern(B, n, p, cones, _sing_mask(n, 22, (18, 19)), 0.1, seed0=999)
    pick = lambda a, b: np.ascontiguousarray(np.where((np.arange(B) % 2 == 0).reshape((B,) + (1,) * (a.ndim - 1)), a, b))
    prob = sb.BatchProblem(pick(reg.c, sng.c), pick(reg.A_cm, sng.A_cm), pick(reg.b, sng.b), pick(reg.G_cm, sng.G_cm),
                           pick(reg.h, sng.h), cones, sing=None, colmajor=True)
    ss = sb.SolverState(prob)
    res = sb.solve_socp_batch(prob, ss)                    # reload=True -> socp_b200_solve_host, sing == NULL
    assert res.timings["path_used"] == sb.PATH_FUSED
    flags = (np.arange(B) % 2).astype(np.uint8)
    assert np.array_equal(ss.get_sing(), flags)
    ref = co.solve_batch(prob.c, prob.A_cm, prob.b, prob.G_cm, prob.h, ocones(cones), sing=flags, nthreads=4)
    conv = ref["status"] == sb.STATUS_CONVERGED
    assert conv[flags == 1].mean() >= 0.8                  # the sing problems are solvable (and solved) with A'A
    agree = (res.status == sb.STATUS_CONVERGED) == conv
    assert agree.mean() >= 0.9, (res.status, ref["status"])
    same = agree & conv & (res.iters == ref["iters"])
    assert same.sum() >= B // 2
    rel = lambda a, b: np.abs(a - b) / np.maximum(1.0, np.abs(b))
    d = np.maximum(rel(res.pobj[same], ref["pobj"][same]), rel(res.dobj[same], ref["dobj"][same]))
    assert d.max() <= 1e-5 and np.median(d) <= 1e-7, (d.max(), np.median(d))
    # the same batch with the flags given, and on the tiled path, finds the same flags / outcomes
    given = sb.BatchProblem(prob.c, prob.A_cm, prob.b, prob.G_cm, prob.h, cones, sing=flags, colmajor=True)
    rg = sb.solve_socp_batch(given, sb.SolverState(given))
    assert np.array_equal(rg.status, res.status) and np.array_equal(rg.pobj, res.pobj)
    til = sb.solve_socp_batch(prob, sb.SolverState(prob), sb.default_params(path=sb.PATH_TILED))
    assert ((til.status == sb.STATUS_CONVERGED) == conv).mean() >= 0.9


def test_compressed_pattern_matches_dense_plan():
    """C2's G = [-I; 0'; -F]: the fused kernel keeps rows 0..49 as (column, value) pairs, drops row 50 and stores only
    the 50 dense rows.  SOCP_B200_NO_V3=1 runs the first-generation kernel on the full 101 x 50 matrix: same statuses and
    iteration counts, objectives within 1e-9 (the two differ in summation order only)."""
    import subprocess, sys, json, os
    prob = gen.make_config("C2", batch=600)
    res = sb.solve_socp_batch(prob, sb.SolverState(prob))
    code = ("import sys, json, numpy as np; sys.path[:0] = %r; import socp_b200 as sb; from socp_b200 import generators as gen;"
            "prob = gen.make_config('C2', batch=600); r = sb.solve_socp_batch(prob, sb.SolverState(prob));"
            "print(json.dumps(dict(status=r.status.tolist(), iters=r.iters.tolist(), pobj=r.pobj.tolist())))") % (sys.path[:4],)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, SOCP_B200_NO_V3="1"))
    assert out.returncode == 0, out.stderr[-2000:]
    old = json.loads(out.stdout.strip().splitlines()[-1])
    assert np.array_equal(res.status, np.array(old["status"])) and (res.status == sb.STATUS_CONVERGED).all()
    assert np.array_equal(res.iters, np.array(old["iters"]))
    assert np.max(np.abs(res.pobj - np.array(old["pobj"])) / np.maximum(1.0, np.abs(res.pobj))) <= 1e-9


def test_pattern_violation_in_a_later_chunk_falls_back():
    """The pipelined one-shot solve takes the row pattern of G from its first chunk; a problem of a later chunk with a
    nonzero outside that pattern is reported by the kernel (verify mode) and the shard is solved again on the all-dense
    plan -- the results must be those of the same batch loaded with set_data (pattern over the whole batch)."""
    prob = gen.make_config("C2", batch=5000)          # > 8 waves of 592 resident CTAs: several chunks
    G = prob.G_cm.copy()
    G[4100, 7, 3] = 0.25                               # row 3 of problem 4100 gets a second nonzero (column 7)
    bad = sb.BatchProblem(prob.c, prob.A_cm, prob.b, G, prob.h, prob.cones, sing=prob.sing, colmajor=True)
    one = sb.solve_socp_batch(bad, sb.SolverState(bad))                      # solve_host, pipelined
    ss2 = sb.SolverState(bad)
    ss2.load(bad)
    two = sb.solve_socp_batch(bad, ss2, reload=False)                        # set_data + solve
    assert (one.status == sb.STATUS_CONVERGED).all() and np.array_equal(one.status, two.status)
    assert np.all(np.abs(one.iters.astype(int) - two.iters.astype(int)) <= 1) and (one.iters == two.iters).mean() >= 0.99
    assert np.max(np.abs(one.pobj - two.pobj) / np.maximum(1.0, np.abs(two.pobj))) <= 1e-9
    from oracle import socp_oracle as so
    pr = so.Problem.create(bad.c[4100], bad.A_dense(4100), bad.b[4100], bad.G_dense(4100), bad.h[4100], ocones(bad.cones), sing=False)
    ref = so.solve_socp(pr, init="reduced", fast_iprod=True)
    assert ref.status == one.status[4100] and abs(ref.iters - int(one.iters[4100])) <= 1
    assert abs(ref.pobj - one.pobj[4100]) <= 1e-8 * max(1.0, abs(ref.pobj))


# ---------------------------------------------------------------- SqrScaling vectors (src/sqrscalings.jl)
def _packed_to_dense(cones, D, u, v):
    """diag(D) + sum_c (u_c u_c' - v_c v_c') from the packed [k] vectors (src/sqrscalings.jl:196-214)."""
    k = so.total_dim(cones)
    out = np.diag(np.asarray(D, dtype=np.float64))
    for kind, offs, dim in cones:
        if kind == so.SOC:
            uc, vc = np.zeros(k), np.zeros(k)
            uc[offs:offs + dim] = u[offs:offs + dim]
            vc[offs:offs + dim] = v[offs:offs + dim]
            out += np.outer(uc, uc) - np.outer(vc, vc)
    return out


@pytest.mark.parametrize("ocn,s,z", [(rc.NT_CONES, rc.NT_S, rc.NT_Z), (rc.SQ_CONES, rc.SQ_U1, rc.SQ_V1),
                                     (rc.SQ_CONES, rc.SQ_U2, rc.SQ_V2)])
def test_sqr_scaling_golden(ocn, s, z):
    """The reference's own pins of the SqrScaling form, test/runtests.jl:51-90: D + uu' - vv' == iWiW, and the reduced
    matrix G'(D + uu' - vv')G == G' iWiW G (:74-77), with D, u, v from the device."""
    cones = bcones(ocn)
    k = so.total_dim(ocn)
    prob = sb.Problem(np.zeros(1), np.zeros((0, 1)), np.zeros(0), np.zeros((k, 1)), np.zeros(k), cones, sing=False)
    ss = sb.SolverState(prob, sb.B200Solver(prob))
    sc = sb.compute_scaling(cones, ss.scaling, s, z)
    D, u, v = sb.sqr_scaling(cones, sc)
    osc = so.Scaling.create(ocn)
    so.compute_scaling(ocn, osc, s, z)
    o2 = so.SqrScaling.create(ocn)
    so.compute_sqr_scaling(ocn, o2, s, z)
    assert relerr(D[0], o2.iWiW) < 1e-13
    assert relerr(u[0], sum(o2.us)) < 1e-13 and relerr(v[0], sum(o2.vs)) < 1e-13
    full = _packed_to_dense(ocn, D[0], u[0], v[0])
    assert np.linalg.norm(full - osc.iWiW) < 1e-12                                            # :80, :89 (1e-2 there)
    if ocn is rc.SQ_CONES:
        H1, H2 = rc.SQ_G.T @ osc.iWiW @ rc.SQ_G, rc.SQ_G.T @ full @ rc.SQ_G                   # :74-77
        assert np.linalg.norm(H2 - H1) < 1e-12


def test_sqr_scaling_batch_vs_oracle_and_closed_form():
    """A batch on the `mixed` layout (orthant chunks + cones up to 70): D, u, v against the oracle, and
    (diag(D) + uu' - vv') x against the kernel's own closed-form W^-2 x (socp_b200_iwiw)."""
    ocn = LAYOUTS["mixed"]
    cones = bcones(ocn)
    k = so.total_dim(ocn)
    B = 9
    rng = np.random.default_rng(5)
    s, z = interior(ocn, rng, B), interior(ocn, rng, B)
    prob = sb.BatchProblem(np.zeros((B, 1)), np.zeros((B, 0, 1)), np.zeros((B, 0)), np.zeros((B, k, 1)), np.zeros((B, k)),
                           cones, sing=np.zeros(B, dtype=np.uint8))
    ss = sb.SolverState(prob)
    sc = sb.compute_scaling(cones, ss.scaling, s, z)
    D, u, v = sb.sqr_scaling(cones, sc)
    xin = rng.standard_normal((B, k))
    ref = np.zeros((B, k))
    sb.iwiw(cones, sc, xin, ref)
    for b in range(B):
        o2 = so.SqrScaling.create(ocn)
        so.compute_sqr_scaling(ocn, o2, s[b], z[b])
        assert relerr(D[b], o2.iWiW) < 1e-12 and relerr(u[b], sum(o2.us)) < 1e-12 and relerr(v[b], sum(o2.vs)) < 1e-12
        assert relerr(_packed_to_dense(ocn, D[b], u[b], v[b]) @ xin[b], ref[b]) < 1e-11


# ---------------------------------------------------------------- lane-per-problem kernel (fused_lane.cuh)
def _with_env(name, value, fn):
    import os
    old = os.environ.get(name)
    os.environ[name] = value
    try:
        return fn()
    finally:
        if old is None:
            del os.environ[name]
        else:
            os.environ[name] = old


@pytest.mark.parametrize("lpw", ["8", "16", "32"])
def test_lane_kernel_c3_vs_oracle_and_one_warp_kernel(lpw):
    """C3 layout, 3000 problems (~20 per SM: lanes take problems one after the other through the work queue): the
    lane-per-problem kernel against the numpy oracle on a sample and against fused_v2's one-warp teams on everything."""
    prob = gen.make_config("C3", batch=3000)
    run = lambda: sb.solve_socp_batch(prob, sb.SolverState(prob))
    res = _with_env("SOCP_B200_LANE", "1", lambda: _with_env("SOCP_B200_LANE_LPW", lpw, run))
    old = _with_env("SOCP_B200_LANE", "0", run)
    assert res.timings["path_used"] == sb.PATH_FUSED
    assert (res.status == sb.STATUS_CONVERGED).all()
    _check_batch(prob, res, range(0, 3000, 100))
    assert np.array_equal(res.status, old.status)
    assert np.all(np.abs(res.iters.astype(int) - old.iters.astype(int)) <= 1)
    same = res.iters == old.iters
    assert same.mean() > 0.99
    rel = lambda a, b: np.abs(a - b) / np.maximum(1.0, np.abs(b))
    d = np.maximum(rel(res.pobj[same], old.pobj[same]), rel(res.dobj[same], old.dobj[same]))
    assert d.max() <= 1e-7 and (d <= 1e-8).mean() >= 0.99, (d.max(), (d <= 1e-8).mean())


def test_lane_kernel_orthant_block_small_batch_and_solve_host():
    """The second instantiation (orthant block + three SOC(3)), a batch smaller than one CTA's lanes, and the
    pipelined one-shot entry (chunks on alternating streams use alternating workspace sets)."""
    cones = [sb.POC(0, 5)] + [sb.SOC(5 + 3 * i, 3) for i in range(3)]
    prob = gen.random_feasible(37, 6, 0, cones, 0.3, 0, 11)
    res = _with_env("SOCP_B200_LANE", "1", lambda: sb.solve_socp_batch(prob, sb.SolverState(prob)))
    old = _with_env("SOCP_B200_LANE", "0", lambda: sb.solve_socp_batch(prob, sb.SolverState(prob)))
    assert np.array_equal(res.status, old.status)
    same = res.iters == old.iters
    assert same.sum() >= 33
    assert np.max(np.abs(res.pobj[same] - old.pobj[same]) / np.maximum(1.0, np.abs(old.pobj[same]))) <= 1e-6
    big = gen.make_config("C3", batch=120000)                          # > 8 waves of 148 x 96 lanes: three chunks
    one_shot = sb.solve_socp_batch(big, sb.SolverState(big))          # socp_b200_solve_host
    assert one_shot.timings["kernel_launches"] >= 2
    ss2 = sb.SolverState(big)
    ss2.load(big)
    two_step = sb.solve_socp_batch(big, ss2, reload=False)             # set_data + solve: one launch
    for key in ("status", "iters", "pobj", "dobj", "x", "z", "s"):
        assert np.array_equal(getattr(one_shot, key), getattr(two_step, key)), key


def test_fused_panel_steps_small_batch_large_n():
    """A handful of n = 500 problems (n not a multiple of the 64-wide panels, a partial last row block): the blocked
    Cholesky takes its panel steps through k_panel_fused (pending update + diagonal block + panel solve in one launch,
    row CTAs released by a flag).  Whole solves against the numpy oracle; the factor against L L' = H at step level."""
    prob = gen.make_config("C4", batch=6)
    ss = sb.SolverState(prob)
    res = sb.solve_socp_batch(prob, ss, sb.default_params(path=sb.PATH_TILED))
    assert (res.status == sb.STATUS_CONVERGED).all()
    _check_batch(prob, res, [0, 5])
    # step level: factor at a strictly interior scaling point, then solve_kkt against the oracle's DenseSolver
    rng = np.random.default_rng(3)
    oc_ = ocones(prob.cones)
    s, z = interior(oc_, rng, prob.B), interior(oc_, rng, prob.B)
    solver = sb.B200Solver(prob)
    ss2 = sb.SolverState(prob, solver)
    ss2.load(prob)
    sc = sb.compute_scaling(prob.cones, ss2.scaling, s, z)
    assert not sb.setup_iter(solver, prob, None, sc).any()
    n, k, B = prob.n, prob.k, prob.B
    dx, dz, ds = rng.standard_normal((B, n)), rng.standard_normal((B, k)), rng.standard_normal((B, k))
    cx, cy, cz, cs = np.zeros((B, n)), np.zeros((B, 0)), np.zeros((B, k)), np.zeros((B, k))
    sb.solve_kkt(solver, prob, None, sc, dx, np.zeros((B, 0)), dz, ds, cx, cy, cz, cs)
    for q in (1, 4):
        pr = so.Problem.create(prob.c[q], prob.A_dense(q), prob.b[q], prob.G_dense(q), prob.h[q], oc_, sing=False)
        osc = so.Scaling.create(oc_)
        so.compute_scaling(oc_, osc, s[q], z[q])
        dsv = so.DenseSolver(pr)
        dsv.setup_iter(pr, osc)
        ox, oy, oz, os_ = dsv.solve_kkt(pr, osc, dx[q], np.zeros(0), dz[q], ds[q], fast_iprod=True)
        assert relerr(cx[q], ox) < 1e-9 and relerr(cz[q], oz) < 1e-9 and relerr(cs[q], os_) < 1e-9


def test_lane_kernel_specialised_at_run_time():
    """A tiny layout without a compile-time instantiation (n = 8, two orthant rows + four SOC(3) + two SOC(5)): the lane-per-problem
    kernel is specialised for it with NVRTC on first use (csrc/lane_jit.cu) -- against the numpy oracle on a sample and
    against the one-warp teams on the whole batch; a second handle with the same layout reuses the compiled kernel."""
    cones = [sb.POC(0, 2)] + [sb.SOC(2 + 3 * i, 3) for i in range(4)] + [sb.SOC(14, 5), sb.SOC(19, 5)]
    prob = gen.random_feasible(2500, 8, 0, cones, 0.3, 0, 21)
    run = lambda: sb.solve_socp_batch(prob, sb.SolverState(prob))
    res = _with_env("SOCP_B200_LANE", "1", lambda: _with_env("SOCP_B200_JIT_VERBOSE", "1", run))
    old = _with_env("SOCP_B200_LANE", "0", run)
    again = _with_env("SOCP_B200_LANE", "1", run)
    assert res.timings["path_used"] == sb.PATH_FUSED and res.timings["kernel_launches"] >= 1
    assert np.array_equal(res.status, old.status)
    for key in ("status", "iters", "pobj", "dobj", "x"):
        assert np.array_equal(getattr(res, key), getattr(again, key)), key
    conv = res.status == sb.STATUS_CONVERGED
    assert conv.mean() > 0.9
    assert np.all(np.abs(res.iters.astype(int) - old.iters.astype(int))[conv] <= 1)
    same = conv & (res.iters == old.iters)
    rel = lambda a, b: np.abs(a - b) / np.maximum(1.0, np.abs(b))
    d = np.maximum(rel(res.pobj[same], old.pobj[same]), rel(res.dobj[same], old.dobj[same]))
    assert same.mean() > 0.9 and np.quantile(d, 0.99) <= 1e-6, (same.mean(), d.max())
    _check_batch(prob, res, [q for q in range(0, 2500, 250) if conv[q]][:6])


def test_fused3_specialised_at_run_time():
    """A C2-like layout that is not BASELINE.json's (portfolio SOCPs with n = 40): the whole-solve kernel gets the layout
    as compile-time constants through NVRTC (csrc/lane_jit.cu) -- against the numpy oracle on a sample and against the
    runtime-dimension instantiation on the whole batch."""
    prob = gen.portfolio(600, 40)
    run = lambda: sb.solve_socp_batch(prob, sb.SolverState(prob))
    res = _with_env("SOCP_B200_JIT_VERBOSE", "1", run)
    dyn = _with_env("SOCP_B200_NO_F3_JIT", "1", run)
    assert res.timings["path_used"] == sb.PATH_FUSED
    assert (res.status == sb.STATUS_CONVERGED).all()
    assert np.array_equal(res.status, dyn.status)
    assert np.all(np.abs(res.iters.astype(int) - dyn.iters.astype(int)) <= 1)
    same = res.iters == dyn.iters
    assert same.mean() > 0.99
    assert np.max(np.abs(res.pobj[same] - dyn.pobj[same]) / np.maximum(1.0, np.abs(dyn.pobj[same]))) <= 1e-8
    _check_batch(prob, res, range(0, 600, 100))


def test_fused2_specialised_at_run_time():
    """A tiny layout with equality rows (n = 12, p = 3, ten SOC(4): not the lane kernel's family): the one-warp-team
    kernel of fused_v2.cuh gets the layout as compile-time constants through NVRTC -- against its runtime-dimension
    instantiation on the whole batch (the two differ in rounding only: a handful of the ~0.02 % of problems that do not
    converge may change status) and against the numpy oracle on a sample."""
    cones = [sb.SOC(4 * i, 4) for i in range(10)]
    prob = gen.random_feasible(2000, 12, 3, cones, 0.3, 0, 31)
    run = lambda: sb.solve_socp_batch(prob, sb.SolverState(prob))
    res = _with_env("SOCP_B200_JIT_VERBOSE", "1", run)
    dyn = _with_env("SOCP_B200_NO_F2_JIT", "1", run)
    assert res.timings["path_used"] == sb.PATH_FUSED
    assert (res.status == dyn.status).mean() >= 0.998
    conv = (res.status == sb.STATUS_CONVERGED) & (dyn.status == sb.STATUS_CONVERGED)
    assert conv.mean() > 0.99
    assert np.all(np.abs(res.iters.astype(int) - dyn.iters.astype(int))[conv] <= 1)
    same = conv & (res.iters == dyn.iters)
    assert same.mean() > 0.98
    assert np.quantile(np.abs(res.pobj[same] - dyn.pobj[same]) / np.maximum(1.0, np.abs(dyn.pobj[same])), 0.99) <= 1e-8
    # the numpy oracle on a sample; objectives to 1e-6 as for the other randomly generated families with equality rows
    # (the two oracles themselves differ by more than 1e-8 on some of these problems, see test_fused_generic_layouts_*)
    for q in [q for q in range(0, 2000, 200) if same[q]][:6]:
        pr = so.Problem.create(prob.c[q], prob.A_dense(q), prob.b[q], prob.G_dense(q), prob.h[q], ocones(prob.cones), sing=False)
        ref = so.solve_socp(pr, init="reduced", fast_iprod=True)
        assert res.status[q] == ref.status and abs(int(res.iters[q]) - ref.iters) <= 1
        if res.iters[q] == ref.iters:
            assert abs(res.pobj[q] - ref.pobj) <= 1e-6 * max(1.0, abs(ref.pobj))
            assert abs(res.dobj[q] - ref.dobj) <= 1e-6 * max(1.0, abs(ref.dobj))
