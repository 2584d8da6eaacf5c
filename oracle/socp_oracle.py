"""CPU oracle (numpy) for the dense KKT hot path of BenChung/Socp.jl.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (``socp.jl_b200/``) may
import this module; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and only as the
checker / the timed CPU baseline.

What this is: a restatement, function by function, of the reference's Julia code
for the path named in SURVEY.md section 8 -- the Mehrotra loop of
``src/solver.jl:40-152``, the *repaired* ``DenseSolver`` of
``src/densesolver.jl:41-90`` (the committed file uses the undefined names
``et``, ``ss.ALi``, ``At`` and ``ss.issng``; the repair replaces them with
``ss.eyetgt``, ``ss.AtLi``, ``pr.A'`` and the type parameter ``sing`` -- the same
algebra as the tested twin ``src/spsolver.jl:86-130``), the NT scalings of
``src/scalings.jl`` / ``src/sqrscalings.jl`` and the Jordan-algebra helpers of
``src/vectors.jl`` / ``src/mats.jl``.

Parity pinning: Julia is not installed in this image, so the reference itself
cannot be executed here.  The oracle is pinned against every golden vector the
reference's own test file holds for this path (``test/runtests.jl:10-48``,
``:50-93``, ``:95-128``, ``:130-191``) in ``tests/test_oracle_golden.py``.  The
*dense-solver-specific* rounding behaviour is "parity unpinned": the reference
never executes its DenseSolver (``test/runtests.jl:143-145,188-190`` build the
dense state and then solve with the sparse one), so the only pins are
solver-agnostic known answers.

Third-party arithmetic restated here (none of it lives under /root/reference):
LAPACK ``dpotrf``/``dpotrs`` and BLAS ``gemm``/``gemv`` through Julia's
LinearAlgebra stdlib (``src/densesolver.jl:42-51,66-86``, ``src/scalings.jl:108``)
and Julia's sparse ``\\`` for the symmetric-indefinite initial system
(``src/solver.jl:84``); numpy's LAPACK-backed ``cholesky``/``solve`` stand in.

Conventions: cones are a sequence of ``(kind, offs, dim)`` with kind 0 = POC,
1 = SOC and ``offs`` the 0-based offset (= ``Cone.offs``, ``src/Socp.jl:8-16``).
All vectors are float64 numpy arrays; nothing here is vectorised over problems.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

POC = 0
SOC = 1

STATUS_CONVERGED = 0   # stop test of src/solver.jl:122 satisfied
STATUS_MAXITER = 1     # the fixed 40-trip loop of src/solver.jl:105 ran out
STATUS_NUMERICAL = 2   # where the reference would throw (DomainError / PosDefException)

Cone = Tuple[int, int, int]


class NumericalFailure(Exception):
    """Stands for Julia's DomainError (sqrt of a negative, src/scalings.jl:46-47)
    and PosDefException (cholesky!, src/densesolver.jl:47,51)."""


def _sqrt(v: float) -> float:
    # Julia's sqrt throws DomainError for negative reals and returns NaN for NaN.
    # The result is a numpy float64 so that a later 1/0 gives Inf as in Julia
    # instead of Python's ZeroDivisionError.
    if v < 0.0:
        raise NumericalFailure("sqrt of negative")
    return np.float64(math.sqrt(v))


def _cholesky(M: np.ndarray) -> np.ndarray:
    """cholesky!(Hermitian(M)) -> lower factor; PosDefException -> NumericalFailure."""
    if M.shape[0] == 0:
        return np.zeros((0, 0))
    if not np.all(np.isfinite(M)):
        raise NumericalFailure("non-finite matrix in cholesky")
    try:
        return np.linalg.cholesky(M)
    except np.linalg.LinAlgError as e:
        raise NumericalFailure(str(e))


def _chol_solve(L: np.ndarray, B: np.ndarray) -> np.ndarray:
    from scipy.linalg import solve_triangular
    if L.shape[0] == 0:
        return np.zeros_like(B)
    Y = solve_triangular(L, B, lower=True)
    return solve_triangular(L.T, Y, lower=False)


# ----------------------------------------------------------------------------
# cone layout helpers                                     src/Socp.jl:8-18
# ----------------------------------------------------------------------------
def conedim(cone: Cone) -> int:
    return cone[2]


def total_dim(cones: Sequence[Cone]) -> int:
    return cones[-1][1] + cones[-1][2] if len(cones) else 0


# ----------------------------------------------------------------------------
# src/vectors.jl
# ----------------------------------------------------------------------------
def make_e(cones: Sequence[Cone]) -> np.ndarray:
    """Identity element e.  src/vectors.jl:7-24."""
    r = np.zeros(total_dim(cones))
    for kind, offs, dim in cones:
        if kind == POC:
            r[offs:offs + dim] = 1.0
        else:
            r[offs] = 1.0
            r[offs + 1:offs + dim] = 0.0
    return r


def vprod(cones: Sequence[Cone], u: np.ndarray, v: np.ndarray) -> np.ndarray:
    """Jordan product u o v.  src/vectors.jl:58-81."""
    t = np.zeros(total_dim(cones))
    for kind, offs, dim in cones:
        if kind == POC:
            for i in range(offs, offs + dim):
                t[i] = u[i] * v[i]
        else:
            acc = 0.0
            for i in range(offs, offs + dim):
                acc += u[i] * v[i]
            t[offs] = acc
            iu, iv = u[offs], v[offs]
            for i in range(offs + 1, offs + dim):
                t[i] = iu * v[i] + iv * u[i]
    return t


def iprod(cones: Sequence[Cone], lam: np.ndarray, v: np.ndarray) -> np.ndarray:
    """Jordan inverse lam \\ v (arrow-matrix solve).  src/vectors.jl:99-131.
    The SOC branch keeps the reference's O(d^2) double loop, term for term."""
    t = np.zeros(total_dim(cones))
    for kind, offs, dim in cones:
        if kind == POC:
            for i in range(offs, offs + dim):
                t[i] = v[i] / lam[i]
        else:
            i1 = offs
            l1 = lam[i1]
            a = l1 ** 2
            for i in range(offs + 1, offs + dim):
                a -= lam[i] * lam[i]
            t[i1] += v[i1] * l1 / a
            for j in range(offs + 1, offs + dim):
                t[i1] -= v[j] * lam[j] / a
            for i in range(offs + 1, offs + dim):
                t[i] -= v[i1] * lam[i] / a
                for j in range(offs + 1, offs + dim):
                    t[i] += v[j] * ((a if i == j else 0.0) + lam[i] * lam[j]) / (l1 * a)
    return t


def iprod_fast(cones: Sequence[Cone], lam: np.ndarray, v: np.ndarray) -> np.ndarray:
    """O(d) form of iprod; algebraically identical to the
    double loop of src/vectors.jl:105-125.  Used for large cones."""
    t = np.zeros(total_dim(cones))
    for kind, offs, dim in cones:
        if kind == POC:
            t[offs:offs + dim] = v[offs:offs + dim] / lam[offs:offs + dim]
        else:
            l0, l1 = lam[offs], lam[offs + 1:offs + dim]
            v0, v1 = v[offs], v[offs + 1:offs + dim]
            a = l0 * l0 - float(l1 @ l1)
            beta = float(l1 @ v1)
            t[offs] = (l0 * v0 - beta) / a
            t[offs + 1:offs + dim] = (-v0 * l1 + (a * v1 + beta * l1) / l0) / a
    return t


def cgt(cones: Sequence[Cone], x: np.ndarray, dx: np.ndarray) -> bool:
    """x + dx in K (test helper).  src/vectors.jl:136-161."""
    for kind, offs, dim in cones:
        if kind == POC:
            for i in range(offs, offs + dim):
                if x[i] + dx[i] < 0:
                    return False
        else:
            tot = 0.0
            for i in range(offs + 1, offs + dim):
                val = x[i] + dx[i]
                tot += val * val
            if not (math.sqrt(tot) <= x[offs] + dx[offs]):
                return False
    return True


def deg(cones: Sequence[Cone]) -> int:
    """src/vectors.jl:165-179: d per POC block, 1 per SOC."""
    return sum(dim if kind == POC else 1 for kind, _, dim in cones)


# ----------------------------------------------------------------------------
# src/mats.jl
# ----------------------------------------------------------------------------
def max_step(cones: Sequence[Cone], x: np.ndarray) -> float:
    """min t with x + t e in K.  src/mats.jl:1-28."""
    maxim = -math.inf
    for kind, offs, dim in cones:
        if kind == POC:
            minim = math.inf
            for i in range(offs, offs + dim):
                if x[i] < minim:
                    minim = x[i]
            val = -minim
        else:
            sqnrm = 0.0
            for i in range(offs + 1, offs + dim):
                sqnrm += x[i] ** 2
            val = math.sqrt(sqnrm) - x[offs]
        if val > maxim:
            maxim = val
    return maxim


def scmax(cones: Sequence[Cone], l: np.ndarray, x: np.ndarray) -> float:
    """max-step of the lambda-scaled direction.  src/mats.jl:42-86."""
    mxv = -math.inf
    for kind, offs, dim in cones:
        if kind == POC:
            val = -math.inf
            for i in range(offs, offs + dim):
                q = -x[i] / l[i]
                if q > val:
                    val = q
        else:
            i1 = offs
            ai = l[i1] ** 2
            for ii in range(offs + 1, offs + dim):
                ai -= l[ii] ** 2
            a = 1.0 / _sqrt(ai)
            r1 = a * l[i1] * x[i1]
            for ii in range(offs + 1, offs + dim):
                r1 -= a * l[ii] * x[ii]
            cst = (r1 + x[i1]) / (a * l[i1] + 1.0)
            r2s = 0.0
            for ii in range(offs + 1, offs + dim):
                r2s += (a * (x[ii] - cst * a * l[ii])) ** 2
            val = math.sqrt(r2s) - a * r1
        if val > mxv:
            mxv = val
    return mxv


def compute_step(cones: Sequence[Cone], l: np.ndarray, ds: np.ndarray, dz: np.ndarray) -> float:
    """src/mats.jl:30-40."""
    mxs = scmax(cones, l, ds)
    mxz = scmax(cones, l, dz)
    t = max(mxs, mxz, 0.0)
    if t == 0.0:
        return 1.0
    return min(1.0, 1.0 / t)


# ----------------------------------------------------------------------------
# src/scalings.jl  -- dense NT scaling
# ----------------------------------------------------------------------------
@dataclass
class Scaling:
    """struct Scaling, src/scalings.jl:1-20.  ``dense=False`` skips the k x k
    matrices W, iW, iWiW (the block-diagonal per-cone blocks are kept in
    ``Wb``/``iWb``/``iWiWb`` instead) so large layouts stay tractable; the
    arithmetic per block is the same."""
    k: int
    ncones: int
    dense: bool = True
    W: Optional[np.ndarray] = None
    iW: Optional[np.ndarray] = None
    iWiW: Optional[np.ndarray] = None
    l: np.ndarray = field(default=None)      # lambda
    mu: np.ndarray = field(default=None)     # eta per cone (unused for POC)
    wbs: np.ndarray = field(default=None)    # sqrt(s/z) for POC, w-bar for SOC
    Wb: List[np.ndarray] = field(default_factory=list)
    iWb: List[np.ndarray] = field(default_factory=list)
    iWiWb: List[np.ndarray] = field(default_factory=list)

    @staticmethod
    def create(cones: Sequence[Cone], dense: bool = True) -> "Scaling":
        k = total_dim(cones)
        sc = Scaling(k=k, ncones=len(cones), dense=dense)
        sc.l = np.zeros(k)
        sc.mu = np.zeros(len(cones))
        sc.wbs = np.zeros(k)
        if dense:
            sc.W = np.zeros((k, k))
            sc.iW = np.zeros((k, k))
            sc.iWiW = np.zeros((k, k))
        return sc


def compute_scaling(cones: Sequence[Cone], sc: Scaling, s: np.ndarray, z: np.ndarray) -> Scaling:
    """src/scalings.jl:101-110 (driver), :22-30 (POC), :32-99 (SOC)."""
    sc.Wb, sc.iWb, sc.iWiWb = [], [], []
    for cind, (kind, offs, dim) in enumerate(cones):
        if kind == POC:
            w = np.zeros(dim)
            iw = np.zeros(dim)
            for i in range(dim):
                ii = offs + i
                w[i] = _sqrt(s[ii] / z[ii])
                iw[i] = _sqrt(z[ii] / s[ii])
                sc.l[ii] = _sqrt(s[ii] * z[ii])
                sc.wbs[ii] = _sqrt(s[ii] / z[ii])
            Wj, iWj = np.diag(w), np.diag(iw)
        else:
            sik = np.array(s[offs:offs + dim], dtype=np.float64)
            zik = np.array(z[offs:offs + dim], dtype=np.float64)
            onrmz = zik[0] ** 2
            onrms = sik[0] ** 2
            for i in range(1, dim):
                onrmz -= zik[i] * zik[i]
                onrms -= sik[i] * sik[i]
            nrmz = _sqrt(onrmz)
            nrms = _sqrt(onrms)
            zik *= 1.0 / nrmz
            sik *= 1.0 / nrms
            zbk, sbk = zik, sik
            nsum = 0.0
            for i in range(dim):
                nsum += zbk[i] * sbk[i]
            gamma = _sqrt((1.0 + nsum) / 2.0)
            wb = np.zeros(dim)
            wb[0] = sbk[0] + zbk[0]
            for i in range(1, dim):
                wb[i] = sbk[i] - zbk[i]
            wb *= 1.0 / (2.0 * gamma)
            sc.wbs[offs:offs + dim] = wb
            denom = wb[0] + 1.0
            mu = _sqrt(nrms / nrmz)
            sc.mu[cind] = mu
            Wj = np.zeros((dim, dim))
            iWj = np.zeros((dim, dim))
            cell = np.eye(dim - 1) + np.outer(wb[1:], wb[1:]) / denom
            Wj[1:, 1:] = cell * mu
            iWj[1:, 1:] = cell / mu
            Wj[0, :] = wb * mu
            iWj[0, 0] = wb[0] / mu
            Wj[1:, 0] = wb[1:] * mu
            iWj[0, 1:] = -wb[1:] / mu
            iWj[1:, 0] = -wb[1:] / mu
            ziv, siv = zbk[0], sbk[0]
            tmv1 = _sqrt(nrms * nrmz)
            mult = tmv1 / (ziv + siv + 2.0 * gamma)
            sbk = sbk * (gamma + ziv)
            zbk = zbk * (gamma + siv)
            for i in range(1, dim):
                sc.l[offs + i] = (sbk[i] + zbk[i]) * mult
            sc.l[offs] = gamma * tmv1
        sc.Wb.append(Wj)
        sc.iWb.append(iWj)
        # mul!(scaling.iWiW, scaling.iW, scaling.iW')  src/scalings.jl:108 -- iW is
        # block diagonal, so the product is formed block by block (identical values).
        sc.iWiWb.append(iWj @ iWj.T)
        if sc.dense:
            sl = slice(offs, offs + dim)
            sc.W[sl, sl] = Wj
            sc.iW[sl, sl] = iWj
            sc.iWiW[sl, sl] = sc.iWiWb[-1]
    return sc


def scale(cones: Sequence[Cone], sc, v: np.ndarray) -> np.ndarray:
    """op = W v, matrix free.  src/scalings.jl:112-117,126-140,159-165."""
    op = np.zeros(total_dim(cones))
    wb = sc.wbs
    for cind, (kind, offs, dim) in enumerate(cones):
        if kind == POC:
            for i in range(offs, offs + dim):
                op[i] = wb[i] * v[i]
        else:
            mu = sc.mu[cind]
            dl = 0.0
            for i in range(offs + 1, offs + dim):
                dl += wb[i] * v[i]
            cst = v[offs] + dl / (1.0 + wb[offs])
            op[offs] = mu * (wb[offs] * v[offs] + dl)
            for i in range(offs + 1, offs + dim):
                op[i] = mu * (v[i] + cst * wb[i])
    return op


def iscale(cones: Sequence[Cone], sc, v: np.ndarray) -> np.ndarray:
    """op = W^-1 v, matrix free.  src/scalings.jl:119-124,142-156,167-173."""
    op = np.zeros(total_dim(cones))
    wb = sc.wbs
    for cind, (kind, offs, dim) in enumerate(cones):
        if kind == POC:
            for i in range(offs, offs + dim):
                op[i] = 1.0 / wb[i] * v[i]
        else:
            mu = sc.mu[cind]
            dl = 0.0
            for i in range(offs + 1, offs + dim):
                dl += wb[i] * v[i]
            cst = -v[offs] + dl / (1.0 + wb[offs])
            op[offs] = 1.0 / mu * (wb[offs] * v[offs] - dl)
            for i in range(offs + 1, offs + dim):
                op[i] = 1.0 / mu * (v[i] + cst * wb[i])
    return op


# ----------------------------------------------------------------------------
# src/sqrscalings.jl -- diagonal + rank-2 form  W^-2 = D + u u' - v v'
# ----------------------------------------------------------------------------
@dataclass
class SqrScaling:
    """struct SqrScaling, src/sqrscalings.jl:8-43 (CHOLMOD workspaces omitted)."""
    iWiW: np.ndarray        # diagonal
    iW: np.ndarray          # diagonal
    l: np.ndarray
    us: List[np.ndarray]
    vs: List[np.ndarray]
    mu: np.ndarray
    wbs: np.ndarray

    @staticmethod
    def create(cones: Sequence[Cone]) -> "SqrScaling":
        k = total_dim(cones)
        return SqrScaling(np.zeros(k), np.zeros(k), np.zeros(k),
                          [np.zeros(k) for _ in cones], [np.zeros(k) for _ in cones],
                          np.zeros(len(cones)), np.zeros(k))


def compute_sqr_scaling(cones: Sequence[Cone], sc: SqrScaling, s: np.ndarray, z: np.ndarray) -> SqrScaling:
    """src/sqrscalings.jl:50-58 (POC), :66-139 (SOC), :177-185 (driver)."""
    for cind, (kind, offs, dim) in enumerate(cones):
        if kind == POC:
            for ii in range(offs, offs + dim):
                sc.iWiW[ii] = z[ii] / s[ii]
                sc.iW[ii] = _sqrt(z[ii] / s[ii])
                sc.l[ii] = _sqrt(s[ii] * z[ii])
                sc.wbs[ii] = _sqrt(s[ii] / z[ii])
            continue
        sbk = np.array(s[offs:offs + dim], dtype=np.float64)
        zbk = np.array(z[offs:offs + dim], dtype=np.float64)

        def cone_prod(vect):
            b = vect[0] * vect[0]
            for i in range(1, dim):
                b -= vect[i] * vect[i]
            return b

        sprod, zprod = cone_prod(sbk), cone_prod(zbk)
        sbk *= 1.0 / _sqrt(sprod)
        zbk *= 1.0 / _sqrt(zprod)
        nsum = 0.0
        for i in range(dim):
            nsum += zbk[i] * sbk[i]
        gamma = _sqrt((1.0 + nsum) / 2.0)
        wb = np.zeros(dim)
        wb[0] = (sbk[0] + zbk[0]) / (2.0 * gamma)
        for i in range(1, dim):
            wb[i] = (sbk[i] - zbk[i]) / (2.0 * gamma)
        mu = _sqrt(_sqrt(sprod / zprod))
        sc.mu[cind] = mu
        inusq = 1.0 / _sqrt(sprod / zprod)
        inu = 1.0 / _sqrt(_sqrt(sprod / zprod))
        wb0 = wb[0]
        wb1 = wb[1:]
        wb1sq = 0.0
        for i in range(dim - 1):
            wb1sq += wb1[i] * wb1[i]
        cv = -(1.0 + wb0 + wb1sq / (1.0 + wb0))
        d = 1.0 + 2.0 / (1.0 + wb0) + wb1sq / ((1.0 + wb0) * (1.0 + wb0))
        a = (wb0 * wb0 + wb1sq - cv * cv * wb1sq / (1.0 + d * wb1sq)) / 2.0
        u0 = _sqrt(wb0 * wb0 + wb1sq - a)
        u1 = cv / u0
        v1 = _sqrt(cv * cv / (u0 * u0) - d)
        sc.iWiW[offs] = a * inusq
        sc.iW[offs] = math.sqrt(abs(a * inusq))
        for i in range(1, dim):
            sc.iWiW[offs + i] = inusq
            sc.iW[offs + i] = math.sqrt(inusq)
        scu, scv = sc.us[cind], sc.vs[cind]
        scu[:] = 0.0
        scv[:] = 0.0
        scu[offs] = inu * u0
        scv[offs] = 0.0
        for i in range(1, dim):
            wbv = inu * wb1[i - 1]
            scu[offs + i] = u1 * wbv
            scv[offs + i] = v1 * wbv
        ziv, siv = zbk[0], sbk[0]
        tmv1 = _sqrt(_sqrt(sprod) * _sqrt(zprod))
        mult = tmv1 / (ziv + siv + 2.0 * gamma)
        for i in range(1, dim):
            sc.l[offs + i] = (sbk[i] * (gamma + ziv) + zbk[i] * (gamma + siv)) * mult
        sc.wbs[offs:offs + dim] = wb
        sc.l[offs] = gamma * tmv1
    return sc


def compute_full_scaling(cones: Sequence[Cone], sc: SqrScaling) -> np.ndarray:
    """Materialise D + u u' - v v' (test helper).  src/sqrscalings.jl:196-214."""
    k = total_dim(cones)
    out = np.zeros((k, k))
    for cind, (kind, offs, dim) in enumerate(cones):
        sl = slice(offs, offs + dim)
        out[sl, sl] = np.diag(sc.iWiW[sl])
        if kind == SOC:
            out += np.outer(sc.us[cind], sc.us[cind])
            out -= np.outer(sc.vs[cind], sc.vs[cind])
    return out


# ----------------------------------------------------------------------------
# src/Socp.jl  -- problem model
# ----------------------------------------------------------------------------
def detect_sing(G: np.ndarray) -> bool:
    """``sing`` of Problem{C,n,m,k,sing}: cholesky(G'G) throws.  src/Socp.jl:49-56."""
    n = G.shape[1]
    if n == 0:
        return False
    try:
        _cholesky(G.T @ G)
        return False
    except NumericalFailure:
        return True


@dataclass
class Problem:
    """struct Problem, src/Socp.jl:20-60 (A and G held dense here)."""
    c: np.ndarray
    A: np.ndarray
    b: np.ndarray
    G: np.ndarray
    h: np.ndarray
    cones: Sequence[Cone]
    n: int = 0
    m: int = 0
    k: int = 0
    sing: bool = False

    @staticmethod
    def create(c, A, b, G, h, cones, sing: Optional[bool] = None) -> "Problem":
        c = np.asarray(c, dtype=np.float64)
        n = c.shape[0]
        A = np.asarray(A, dtype=np.float64).reshape(-1, n)
        b = np.asarray(b, dtype=np.float64).reshape(-1)
        G = np.asarray(G, dtype=np.float64).reshape(-1, n)
        h = np.asarray(h, dtype=np.float64).reshape(-1)
        m, k = A.shape[0], G.shape[0]
        assert b.shape[0] == m and h.shape[0] == k      # src/Socp.jl:43-47
        assert total_dim(cones) == k
        if sing is None:
            sing = detect_sing(G)
        return Problem(c, A, b, G, h, tuple(cones), n, m, k, bool(sing))


@dataclass
class State:
    """struct State, src/Socp.jl:62-75."""
    x: np.ndarray
    y: np.ndarray
    z: np.ndarray
    s: np.ndarray


# ----------------------------------------------------------------------------
# src/densesolver.jl (repaired)
# ----------------------------------------------------------------------------
class DenseSolver:
    """mutable struct DenseSolver, src/densesolver.jl:1-39."""

    def __init__(self, pr: Problem):
        self.AA = pr.A.T @ pr.A          # :32
        self.GWiWi = None
        self.L = None                    # GWiWiGfact
        self.Li = None                   # explicit inverse of G'W^-2 G (+A'A)
        self.AtLi = None
        self.LA = None                   # AtLiAfact
        self.H = None

    def setup_iter(self, pr: Problem, sc: Scaling) -> None:
        """KKT factor.  src/densesolver.jl:41-52 with et->eyetgt, ALi->AtLi, At->A'."""
        n = pr.n
        GWiWi = np.zeros((n, pr.k))
        for (kind, offs, dim), blk in zip(pr.cones, sc.iWiWb):
            sl = slice(offs, offs + dim)
            GWiWi[:, sl] = pr.G[sl, :].T @ blk           # :42
        H = GWiWi @ pr.G                                  # :43
        if pr.sing:
            H = H + self.AA                               # :44-46
        self.H = H
        self.GWiWi = GWiWi
        self.L = _cholesky(H)                             # :47
        self.Li = _chol_solve(self.L, np.eye(n))          # :48
        self.AtLi = pr.A @ self.Li                        # :49
        AtLiA = self.AtLi @ pr.A.T                        # :50
        self.LA = _cholesky(AtLiA)                        # :51

    def solve_kkt(self, pr: Problem, sc: Scaling, dx, dy, dz, ds, fast_iprod: bool = False):
        """KKT solve.  src/densesolver.jl:54-90 with ss.issng -> sing."""
        ip = iprod_fast if fast_iprod else iprod
        k0 = ip(pr.cones, sc.l, ds)                       # :61
        k1 = scale(pr.cones, sc, k0)                      # :62
        k2 = dz - k1                                      # :65
        n0 = self.GWiWi @ k2                              # :66
        n0 = n0 + dx                                      # :67
        if pr.sing:
            n0 = n0 + pr.A.T @ dy                         # :69-71
        m0 = self.AtLi @ n0                               # :73
        m0 = m0 - dy                                      # :74
        cy = _chol_solve(self.LA, m0)                     # :75
        if pr.sing:
            m0 = dy - cy                                  # :76-77
        else:
            m0 = -cy                                      # :78-79
        n1 = pr.A.T @ m0                                  # :81
        n0 = n0 + n1                                      # :82
        cx = self.Li @ n0                                 # :83
        k1 = pr.G @ cx                                    # :84
        k1 = k1 - k2                                      # :85
        cz = np.zeros(pr.k)
        for (kind, offs, dim), blk in zip(pr.cones, sc.iWiWb):
            sl = slice(offs, offs + dim)
            cz[sl] = blk @ k1[sl]                         # :86
        k1 = scale(pr.cones, sc, cz)                      # :87
        k0 = k0 - k1                                      # :88
        cs = scale(pr.cones, sc, k0)                      # :89
        return cx, cy, cz, cs


# ----------------------------------------------------------------------------
# src/solver.jl
# ----------------------------------------------------------------------------
@dataclass
class Params:
    """The reference's literals: 40 trips (src/solver.jl:105), 1e-5 (:122),
    0.99 (:146), 1e-10 (:91,:97)."""
    max_iter: int = 40
    tol: float = 1e-5
    step_damp: float = 0.99
    init_eps: float = 1e-10


@dataclass
class Result:
    state: State
    status: int
    iters: int
    pobj: float
    dobj: float
    trace: list = field(default_factory=list)


def initial_point_full(pr: Problem) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Literal restatement of src/solver.jl:68-84: assemble the (n+m+k) symmetric
    indefinite matrix [0 A' G'; A 0 0; G 0 -I] and solve against [-c; b; h]."""
    n, m, k = pr.n, pr.m, pr.k
    M = np.zeros((n + m + k, n + m + k))
    M[:n, n:n + m] = pr.A.T
    M[:n, n + m:] = pr.G.T
    M[n:n + m, :n] = pr.A
    M[n + m:, :n] = pr.G
    M[n + m:, n + m:] = -np.eye(k)
    rhs = np.concatenate([-pr.c, pr.b, pr.h])
    sol = np.linalg.solve(M, rhs)
    return sol[:n], sol[n:n + m], sol[n + m:]


def initial_point_reduced(pr: Problem) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """The same system by block elimination with W = I:
    H = G'G (+A'A if sing), n0 = -c + G'h (+A'b), M = A H^-1 A',
    y = M^-1 (A H^-1 n0 - b), x = H^-1 (n0 - A'y), z = G x - h.
    With sing the substitution is the *correct* one (x = H^-1(n0 - A'y) with the
    regularised H solves (G'G + A'A) x + A'(y - b)... see note below), not the
    loop-only quirk of src/densesolver.jl:76-82."""
    H = pr.G.T @ pr.G
    n0 = -pr.c + pr.G.T @ pr.h
    if pr.sing:
        H = H + pr.A.T @ pr.A
        n0 = n0 + pr.A.T @ pr.b
    L = _cholesky(H)
    if pr.m > 0:
        HiAt = _chol_solve(L, pr.A.T)
        M = pr.A @ HiAt
        LA = _cholesky(M)
        t = _chol_solve(L, n0)
        # with sing, H x + A' y~ = n0 where y~ = y - b*0 ... the regularised system is
        # (G'G + A'A) x + A' y = -c + G'h + A'b  <=>  G'G x + A'(y + A x - b) = ...,
        # and A x = b at the solution, so y is unchanged.
        y = _chol_solve(LA, pr.A @ t - pr.b)
        x = t - HiAt @ y
    else:
        y = np.zeros(0)
        x = _chol_solve(L, n0)
    z = pr.G @ x - pr.h
    return x, y, z


def solve_socp(pr: Problem, params: Params = Params(), init: str = "full",
               fast_iprod: bool = False, dense_scaling: bool = True,
               trace: bool = False) -> Result:
    """solve_socp, src/solver.jl:40-152, with the DenseSolver back end.
    ``init``: "full" = literal (n+m+k) indefinite solve (:68-84); "reduced" = the
    block-eliminated form the GPU path uses."""
    cones = pr.cones
    n, m, k = pr.n, pr.m, pr.k
    try:
        if init == "full":
            x0, y0, iz = initial_point_full(pr)
        else:
            x0, y0, iz = initial_point_reduced(pr)
    except (NumericalFailure, np.linalg.LinAlgError):
        z = np.zeros(k)
        st = State(np.zeros(n), np.zeros(m), z, z.copy())
        return Result(st, STATUS_NUMERICAL, 0, 0.0, 0.0)
    idel = make_e(cones)                                   # :86
    alphp = max_step(cones, -iz)                           # :88
    alphd = max_step(cones, iz)                            # :89
    if abs(alphp) < params.init_eps:                       # :91-95
        inits = -iz
    else:
        inits = -iz + (1.0 + alphp) * idel
    if abs(alphd) < params.init_eps:                       # :97-101
        initz = iz.copy()
    else:
        initz = iz + (1.0 + alphd) * idel
    st = State(x0.copy(), y0.copy(), initz, inits)         # :104
    sc = Scaling.create(cones, dense=dense_scaling)
    solver = DenseSolver(pr)
    status = STATUS_MAXITER
    iters = 0
    tr = []
    dg = deg(cones)
    for it in range(params.max_iter):                      # :105
        try:
          with np.errstate(all="ignore"):                  # Julia: Inf/NaN propagate silently
              compute_scaling(cones, sc, st.s, st.z)         # :106
              l = sc.l
              dx = pr.A.T @ st.y + pr.G.T @ st.z + pr.c      # :110-112
              dy = pr.A @ st.x - pr.b                        # :114-115
              dz = pr.G @ st.x + st.s - pr.h                 # :117-118
              ds = vprod(cones, l, l)                        # :120
              gap = float(st.z @ st.s)
              resid = float(np.linalg.norm(dx) + np.linalg.norm(dy) + gap)
              if trace:
                  tr.append(dict(it=it, resid=resid, gap=gap))
              if resid < params.tol:                         # :122-124
                  status = STATUS_CONVERGED
                  break
              dx, dy, dz, ds = -dx, -dy, -dz, -ds            # :125
              solver.setup_iter(pr, sc)                      # :126
              rx, ry, rz, rs = solver.solve_kkt(pr, sc, dx, dy, dz, ds, fast_iprod)   # :127
              kt3 = scale(cones, sc, rz)                     # :128
              kt2 = iscale(cones, sc, rs)                    # :129
              t = compute_step(cones, l, kt3, kt2)           # :130
              ll = float(l @ l)
              rho = 1.0 - t - t ** 2 * float(kt2 @ kt3) / ll  # :132  (minus: reference quirk)
              sig = max(0.0, min(1.0, rho)) ** 3             # :133
              mu = ll / dg                                   # :134
              scfact = 1.0 - sig                             # :136
              kt1 = vprod(cones, kt2, kt3)                   # :137
              kt2 = (sig * mu) * idel                        # :138
              ds = ds + (kt2 - kt1)                          # :139
              dx, dy, dz = dx * scfact, dy * scfact, dz * scfact   # :140
              rx, ry, rz, rs = solver.solve_kkt(pr, sc, dx, dy, dz, ds, fast_iprod)   # :141
              kt3 = scale(cones, sc, rz)                     # :143
              kt2 = iscale(cones, sc, rs)                    # :144
              step = compute_step(cones, l, kt3, kt2)        # :145
              step *= params.step_damp                       # :146
              if not (np.all(np.isfinite(rx)) and np.all(np.isfinite(rz)) and np.all(np.isfinite(rs))
                      and np.all(np.isfinite(ry)) and math.isfinite(step)):
                  raise NumericalFailure("non-finite step")
        except NumericalFailure:
            status = STATUS_NUMERICAL
            break
        st.x = st.x + rx * step                            # :147
        st.y = st.y + ry * step                            # :148
        st.z = st.z + rz * step                            # :149
        st.s = st.s + rs * step                            # :150
        iters += 1
    pobj = float(pr.c @ st.x)
    dobj = float(-(pr.b @ st.y) - pr.h @ st.z)
    return Result(st, status, iters, pobj, dobj, tr)
