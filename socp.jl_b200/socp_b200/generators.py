"""Synthetic problem generators for BASELINE.json's configs.

Counter-based RNG (Philox) seeded ``seed0 + problem_index`` so that a problem's
bytes do not depend on the batch size or on which rank generates it (shards of
one batch can be generated independently).  Matrices are produced directly in
the library's column-major layout: G (B, n, k), A (B, n, p).
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import numpy as np

from .api import POC, SOC, BatchProblem, Cone

SEED0 = 1234


def _rng(i: int, seed0: int = SEED0) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(seed0 + int(i)))


def portfolio_cones(n: int) -> Tuple[Cone, ...]:
    return (POC(0, n), SOC(n, n + 1))


def portfolio(batch: int, n: int = 50, first: int = 0, seed0: int = SEED0) -> BatchProblem:
    """C2: max mu'x s.t. 1'x = 1, x >= 0, ||F x|| <= gamma.
    c = -mu, A = 1' (1 x n), b = 1, G = [-I; 0'; -F] ((2n+1) x n), h = [0; gamma; 0],
    cones [POC(0,n), SOC(n,n+1)];  F ~ 0.2 N(0,1)/sqrt(n), mu ~ U(0,0.2), gamma = 0.1."""
    k = 2 * n + 1
    gamma = 0.1
    c = np.empty((batch, n))
    G = np.zeros((batch, n, k))          # column-major per problem: G[b, j, i] = G_b[i, j]
    idx = np.arange(n)
    for q in range(batch):
        r = _rng(first + q, seed0)
        F = 0.2 * r.standard_normal((n, n)) / np.sqrt(n)     # F[i, j]
        mu = r.uniform(0.0, 0.2, n)
        c[q] = -mu
        G[q, idx, idx] = -1.0
        G[q, :, n + 1:] = -F.T
    h = np.zeros((batch, k))
    h[:, n] = gamma
    A = np.ones((batch, n, 1))
    b = np.ones((batch, 1))
    return BatchProblem(c, A, b, G, h, portfolio_cones(n), sing=np.zeros(batch, dtype=np.uint8), colmajor=True)


def soc_cones(ncones: int, dim: int) -> Tuple[Cone, ...]:
    return tuple(SOC(j * dim, dim) for j in range(ncones))


def random_feasible(batch: int, n: int, p: int, cones: Sequence[Cone], scale: float, first: int = 0,
                    seed0: int = SEED0, sing_known: bool = True) -> BatchProblem:
    """Strictly primal-dual feasible random instance (C3, C4, C5 and p > 0 variants):
    G ~ N(0,1)/sqrt(n), A ~ N(0,1)/sqrt(n); interior s0, z0 (POC entries U(0.5,2); SOC
    tail N(0,1), head = ||tail|| + U(0.5,1.5)); x0, y0 ~ N(0,1); s0, z0, x0, y0 all
    multiplied by `scale`; h = G x0 + s0, b = A x0, c = -A'y0 - G'z0."""
    cones = tuple(cones)
    k = sum(cn.dim for cn in cones)
    c = np.empty((batch, n))
    G = np.empty((batch, n, k))
    A = np.empty((batch, n, p))
    b = np.empty((batch, p))
    h = np.empty((batch, k))
    for q in range(batch):
        r = _rng(first + q, seed0)
        Gq = r.standard_normal((n, k)) / np.sqrt(n)          # column-major G_b: Gq[j, i] = G_b[i, j]
        Aq = r.standard_normal((n, p)) / np.sqrt(n)
        s0 = np.empty(k)
        z0 = np.empty(k)
        for cn in cones:
            sl = slice(cn.offs, cn.offs + cn.dim)
            if cn.kind == 0:
                s0[sl] = r.uniform(0.5, 2.0, cn.dim)
                z0[sl] = r.uniform(0.5, 2.0, cn.dim)
            else:
                for v in (s0, z0):
                    tail = r.standard_normal(cn.dim - 1)
                    v[cn.offs + 1:cn.offs + cn.dim] = tail
                    v[cn.offs] = np.linalg.norm(tail) + r.uniform(0.5, 1.5)
        x0 = r.standard_normal(n) * scale
        y0 = r.standard_normal(p) * scale
        s0 *= scale
        z0 *= scale
        G[q] = Gq
        A[q] = Aq
        h[q] = Gq.T @ x0 + s0
        b[q] = Aq.T @ x0
        c[q] = -(Aq @ y0) - Gq @ z0
    sing = np.zeros(batch, dtype=np.uint8) if (sing_known and k >= n) else None
    return BatchProblem(c, A, b, G, h, cones, sing=sing, colmajor=True)


def random_feasible_pattern(batch: int, n: int, p: int, cones: Sequence[Cone], mask: np.ndarray, scale: float = 0.1,
                            first: int = 0, seed0: int = SEED0) -> BatchProblem:
    """random_feasible with a sparsity pattern on G: `mask` (k x n booleans) says which entries are stored -- one
    pattern for the whole batch, values per problem (the reference keeps G as a SparseMatrixCSC, src/Socp.jl:29).
    Stored entries are kept away from zero.  Rows with one stored entry are bounds / single-variable cone rows, empty
    rows are constants; zero columns make G rank deficient (`sing`, src/Socp.jl:49-56)."""
    cones = tuple(cones)
    k = sum(cn.dim for cn in cones)
    assert mask.shape == (k, n)
    c = np.empty((batch, n))
    G = np.zeros((batch, n, k))
    A = np.empty((batch, n, p))
    b = np.empty((batch, p))
    h = np.empty((batch, k))
    for q in range(batch):
        r = _rng(first + q, seed0)
        Gq = (r.standard_normal((n, k)) / np.sqrt(n)) * mask.T
        Gq[mask.T & (np.abs(Gq) < 1e-3)] = 0.5
        Aq = r.standard_normal((n, p)) / np.sqrt(n)
        s0 = np.empty(k)
        z0 = np.empty(k)
        for cn in cones:
            sl = slice(cn.offs, cn.offs + cn.dim)
            if cn.kind == 0:
                s0[sl] = r.uniform(0.5, 2.0, cn.dim)
                z0[sl] = r.uniform(0.5, 2.0, cn.dim)
            else:
                for v in (s0, z0):
                    tail = r.standard_normal(cn.dim - 1)
                    v[cn.offs + 1:cn.offs + cn.dim] = tail
                    v[cn.offs] = np.linalg.norm(tail) + r.uniform(0.5, 1.5)
        x0 = r.standard_normal(n) * scale
        y0 = r.standard_normal(p) * scale
        s0 *= scale
        z0 *= scale
        G[q] = Gq
        A[q] = Aq
        h[q] = Gq.T @ x0 + s0
        b[q] = Aq.T @ x0
        c[q] = -(Aq @ y0) - Gq @ z0
    return BatchProblem(c, A, b, G, h, cones, sing=None, colmajor=True)


# BASELINE.json configs: name -> generator kwargs
CONFIGS: Dict[str, dict] = {
    "C2": dict(kind="portfolio", n=50, batch=10_000),
    "C3": dict(kind="random", n=12, p=0, ncones=10, dim=4, scale=0.1, batch=100_000),
    "C4": dict(kind="random", n=500, p=0, ncones=20, dim=50, scale=0.03, batch=1_000),
    "C5": dict(kind="random", n=4096, p=0, ncones=64, dim=128, scale=0.01, batch=1),
}


def make_config(name: str, batch: int = None, first: int = 0, seed0: int = SEED0) -> BatchProblem:
    cfg = CONFIGS[name]
    B = cfg["batch"] if batch is None else batch
    if cfg["kind"] == "portfolio":
        return portfolio(B, cfg["n"], first, seed0)
    return random_feasible(B, cfg["n"], cfg["p"], soc_cones(cfg["ncones"], cfg["dim"]), cfg["scale"], first, seed0)
