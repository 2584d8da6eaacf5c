"""Host-side mirror of the reference's solver interface, calling libsocp_b200
through ctypes.  Same names, argument meaning and error behaviour as the
reference (/root/reference = BenChung/Socp.jl):

    Problem(c, A, b, G, h, cones)            src/Socp.jl:20-60
    State                                     src/Socp.jl:62-75
    SolverState(prob, solver)                 src/solver.jl:1-38
    solve_socp(prob, ss) -> State             src/solver.jl:40-152
    compute_scaling / scale_ / iscale_        src/scalings.jl:101-173  (scale!, iscale!)
    setup_iter / solve_kkt                    src/densesolver.jl:41-90
    vprod / iprod / make_e / deg              src/vectors.jl
    max_step / compute_step                   src/mats.jl

plus the batch entry the GPU needs (BatchProblem, BatchSolverState,
solve_socp_batch): many independent problems that share one cone layout.
Dimension mismatches raise AssertionError like the reference's @assert's;
library errors raise SocpError.  Nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib as L

STATUS_CONVERGED, STATUS_MAXITER, STATUS_NUMERICAL = 0, 1, 2
PATH_AUTO, PATH_TILED, PATH_FUSED = 0, 1, 2
_POC, _SOC = 0, 1


class SocpError(RuntimeError):
    pass


@dataclass(frozen=True)
class Cone:
    """abstract type Cone{D}; kind 0 = POC, 1 = SOC; offs is 0-based (Cone.offs)."""
    kind: int
    offs: int
    dim: int


def POC(offs: int, dim: int) -> Cone:
    """POC(offs, dim): positive-orthant block, reference src/Socp.jl:9-12."""
    return Cone(_POC, int(offs), int(dim))


def SOC(offs: int, dim: int) -> Cone:
    """SOC(offs, dim): second-order cone, reference src/Socp.jl:13-16."""
    return Cone(_SOC, int(offs), int(dim))


def _as_cones(cones) -> Tuple[Cone, ...]:
    out = []
    for c in cones:
        out.append(c if isinstance(c, Cone) else Cone(int(c[0]), int(c[1]), int(c[2])))
    return tuple(out)


def deg(cones) -> int:
    """reference src/vectors.jl:165-179"""
    return sum(c.dim if c.kind == _POC else 1 for c in _as_cones(cones))


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _dp(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(L.c_double_p)


def _ip(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(L.c_int32_p)


def default_params(**kw) -> L.Params:
    p = L.Params()
    L.load().socp_b200_default_params(C.byref(p))
    for k_, v in kw.items():
        setattr(p, k_, v)
    return p


class _Handle:
    """Owns one socp_handle*."""

    def __init__(self, n: int, p: int, cones: Tuple[Cone, ...], batch: int, devices: Optional[Sequence[int]] = None):
        self.lib = L.load()
        self.n, self.p, self.cones, self.batch = int(n), int(p), cones, int(batch)
        self.k = sum(c.dim for c in cones)
        self.ncones = len(cones)
        kind = np.array([c.kind for c in cones], dtype=np.int32)
        offs = np.array([c.offs for c in cones], dtype=np.int32)
        dim = np.array([c.dim for c in cones], dtype=np.int32)
        lay = L.Layout(self.n, self.p, self.k, self.ncones, _ip(kind), _ip(offs), _ip(dim))
        self.ptr = L.H()
        dev = None
        nd = 0
        if devices is not None:
            dev_arr = np.array(list(devices), dtype=np.int32)
            dev, nd = _ip(dev_arr), len(dev_arr)
        rc = self.lib.socp_b200_create(C.byref(self.ptr), C.byref(lay), self.batch, dev, nd)
        if rc != 0:
            msg = self.lib.socp_b200_last_error(None)
            self.ptr = None
            raise SocpError(f"socp_b200_create failed ({rc}): {msg.decode() if msg else ''}")

    def check(self, rc: int, what: str):
        if rc != 0:
            msg = self.lib.socp_b200_last_error(self.ptr)
            raise SocpError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def close(self):
        if getattr(self, "ptr", None):
            self.lib.socp_b200_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def timings(self) -> dict:
        t = L.Timings()
        self.check(self.lib.socp_b200_timings(self.ptr, C.byref(t)), "timings")
        return {f: getattr(t, f) for f, _ in L.Timings._fields_}


# ----------------------------------------------------------------------- problems
class CscMatrix:
    """A (batch of) SparseMatrixCSC{Float64,Int64}, the reference's storage of A and
    G (src/Socp.jl:25,29): one pattern (colptr, rowval), values ``nzval`` of shape
    (nnz,) -- one matrix shared by the batch -- or (B, nnz).  ``index_base`` is 1
    for arrays taken from Julia, 0 for scipy's.  The library assembles the dense
    operands on the device (socp_b200_set_data_csc); nothing is densified here."""

    def __init__(self, shape, colptr, rowval, nzval, index_base: int = 0):
        self.shape = (int(shape[0]), int(shape[1]))
        self.colptr = np.ascontiguousarray(colptr, dtype=np.int64)
        self.rowval = np.ascontiguousarray(rowval, dtype=np.int64)
        self.nzval = _f64(nzval)
        self.index_base = int(index_base)
        assert self.colptr.shape == (self.shape[1] + 1,)
        assert self.nzval.shape[-1] == self.rowval.shape[0]
        self.shared = (self.nzval.ndim == 1)

    @classmethod
    def from_scipy(cls, m, values=None) -> "CscMatrix":
        """From a scipy.sparse matrix (pattern + values), or its pattern with per-problem ``values`` (B, nnz)."""
        m = m.tocsc()
        m.sort_indices()
        return cls(m.shape, m.indptr, m.indices, m.data if values is None else values, 0)

    @property
    def nnz(self) -> int:
        return int(self.rowval.shape[0])

    def c_struct(self) -> L.Csc:
        return L.Csc(self.nnz, self.colptr.ctypes.data_as(L.c_int64_p), self.rowval.ctypes.data_as(L.c_int64_p),
                     self.nzval.ctypes.data_as(L.c_double_p), self.index_base)


class BatchProblem:
    """B independent problems sharing (n, p, cones).

    Logical shapes: c (B,n), A (B,p,n), b (B,p), G (B,k,n), h (B,k).  With
    ``colmajor=True`` A and G are given already in the library's memory layout
    (column-major per problem): A (B,n,p), G (B,n,k) C-contiguous -- no copy.
    ``shared_A`` / ``shared_G``: one matrix (2-D) shared by the whole batch.
    ``sing``: optional (B,) uint8, the reference's 5th type parameter; computed
    on the device when omitted (reference src/Socp.jl:49-56)."""

    def __init__(self, c, A, b, G, h, cones, sing=None, colmajor: bool = False):
        self.cones = _as_cones(cones)
        self.c = _f64(c)
        assert self.c.ndim == 2, "c must be (B, n)"
        self.B, self.n = self.c.shape
        self.k = sum(cn.dim for cn in self.cones)
        self.h = _f64(h)
        assert self.h.shape == (self.B, self.k)                     # src/Socp.jl:47
        self.G_csc = G if isinstance(G, CscMatrix) else None
        self.A_csc = A if isinstance(A, CscMatrix) else None
        if self.G_csc is not None:
            assert self.G_csc.shape == (self.k, self.n)                 # src/Socp.jl:45-46
            assert self.G_csc.shared or self.G_csc.nzval.shape[0] == self.B
            self.shared_G = self.G_csc.shared
            self.G_cm = None
            G = None
        else:
            G = np.asarray(G, dtype=np.float64)
            self.shared_G = (G.ndim == 2)
        if self.G_csc is not None:
            pass
        elif self.shared_G:
            Gl = G.T if colmajor else G
            assert Gl.shape == (self.k, self.n)                     # src/Socp.jl:45-46
            self.G_cm = _f64(Gl.T)                                  # (n, k) C-order == k x n column-major
        else:
            if colmajor:
                assert G.shape == (self.B, self.n, self.k)
                self.G_cm = _f64(G)
            else:
                assert G.shape == (self.B, self.k, self.n)
                self.G_cm = _f64(G.transpose(0, 2, 1))
        b = _f64(b)
        if b.ndim == 1 and b.size == 0:
            b = b.reshape(self.B, 0)
        self.b = b
        assert self.b.ndim == 2 and self.b.shape[0] == self.B
        self.p = self.b.shape[1]
        assert (self.A_csc is not None) == (self.G_csc is not None) or self.p == 0, \
            "A and G must both be CscMatrix (or both dense) when there are equality rows"
        if self.A_csc is not None and self.p > 0:
            assert self.A_csc.shape == (self.p, self.n)                 # src/Socp.jl:43-44
            assert self.A_csc.shared or self.A_csc.nzval.shape[0] == self.B
            self.shared_A = self.A_csc.shared
            self.A_cm = None
        else:
            self.A_csc = None
            A = np.asarray(A, dtype=np.float64)
            self.shared_A = (A.ndim == 2 and self.p > 0)
        if self.A_csc is not None:
            pass
        elif self.p == 0:
            self.A_cm = np.zeros((self.B, self.n, 0))
        elif self.shared_A:
            Al = A.T if colmajor else A
            assert Al.shape == (self.p, self.n)                     # src/Socp.jl:43-44
            self.A_cm = _f64(Al.T)
        else:
            if colmajor:
                assert A.shape == (self.B, self.n, self.p)
                self.A_cm = _f64(A)
            else:
                assert A.shape == (self.B, self.p, self.n)
                self.A_cm = _f64(A.transpose(0, 2, 1))
        self.sing = None if sing is None else np.ascontiguousarray(sing, dtype=np.uint8).reshape(self.B)
        off = 0
        for cn in self.cones:
            assert cn.offs == off, "cones must tile 0..k-1 contiguously"
            off += cn.dim

    @property
    def m(self) -> int:          # the reference calls the equality-row count m
        return self.p

    def G_dense(self, i: int) -> np.ndarray:
        assert self.G_csc is None, "CSC problems are assembled on the device; there is no host-side dense copy"
        return (self.G_cm if self.shared_G else self.G_cm[i]).T

    def A_dense(self, i: int) -> np.ndarray:
        if self.p == 0:
            return np.zeros((0, self.n))
        assert self.A_csc is None, "CSC problems are assembled on the device; there is no host-side dense copy"
        return (self.A_cm if self.shared_A else self.A_cm[i]).T


class Problem(BatchProblem):
    """Problem(c, A, b, G, h, cones), reference src/Socp.jl:20-60 -- a batch of one."""

    def __init__(self, c, A, b, G, h, cones, sing: Optional[bool] = None):
        c = np.asarray(c, dtype=np.float64).reshape(1, -1)
        n = c.shape[1]
        A = np.asarray(A, dtype=np.float64).reshape(-1, n)
        b = np.asarray(b, dtype=np.float64).reshape(-1)
        G = np.asarray(G, dtype=np.float64)
        h = np.asarray(h, dtype=np.float64).reshape(-1)
        assert b.shape[0] == A.shape[0]                             # src/Socp.jl:43
        assert G.ndim == 2 and G.shape[1] == n                      # src/Socp.jl:45
        assert h.shape[0] == G.shape[0]                             # src/Socp.jl:47
        super().__init__(c, A[None] if A.shape[0] else np.zeros((1, 0, n)), b[None], G[None], h[None], cones,
                         None if sing is None else np.array([1 if sing else 0], dtype=np.uint8))


@dataclass
class State:
    """struct State, reference src/Socp.jl:62-75, plus what the reference lacks:
    status / iters / objectives."""
    x: np.ndarray
    y: np.ndarray
    z: np.ndarray
    s: np.ndarray
    status: int = -1
    iters: int = 0
    pobj: float = float("nan")
    dobj: float = float("nan")


@dataclass
class BatchResult:
    x: np.ndarray
    y: np.ndarray
    z: np.ndarray
    s: np.ndarray
    status: np.ndarray
    iters: np.ndarray
    pobj: np.ndarray
    dobj: np.ndarray
    timings: dict


# ----------------------------------------------------------------------- solver state
class B200Solver:
    """The KKTSolver{B200Scaling} subtype of the plug-in seam (reference
    src/Socp.jl:77-78; DenseSolver at src/densesolver.jl:1-39).  Workspaces live
    on the device inside the handle created by SolverState."""

    def __init__(self, prob: Optional[BatchProblem] = None, devices: Optional[Sequence[int]] = None):
        self.devices = devices
        self.handle: Optional[_Handle] = None


class B200Scaling:
    """AbstractScaling of the plug-in seam: fields l (lambda), wbs, mu as in
    struct Scaling (reference src/scalings.jl:1-20); W/iW/iWiW are never
    materialised -- they are applied matrix-free on the device."""

    def __init__(self, handle: _Handle):
        self.handle = handle
        B, k, nc = handle.batch, handle.k, handle.ncones
        self.l = np.zeros((B, k))
        self.wbs = np.zeros((B, k))
        self.mu = np.zeros((B, nc))
        self.fail = np.zeros(B, dtype=np.int32)


class BatchSolverState:
    """SolverState(prob, solver), reference src/solver.jl:1-38, for a batch.
    Reusable across solves (reference test/runtests.jl:243)."""

    def __init__(self, prob: BatchProblem, solver: Optional[B200Solver] = None, devices: Optional[Sequence[int]] = None):
        self.solver = solver or B200Solver(prob, devices)
        if devices is None:
            devices = self.solver.devices
        self.handle = _Handle(prob.n, prob.p, prob.cones, prob.B, devices)
        self.solver.handle = self.handle
        self.scaling = B200Scaling(self.handle)
        self._loaded_id = None

    def load(self, prob: BatchProblem, force: bool = False):
        """Upload the problem data (host -> device)."""
        h = self.handle
        assert (prob.n, prob.p, prob.B) == (h.n, h.p, h.batch) and prob.cones == h.cones
        if not force and self._loaded_id == id(prob):
            return
        flags = (1 if prob.shared_A else 0) | (2 if prob.shared_G else 0)
        sing = None if prob.sing is None else prob.sing.ctypes.data_as(L.c_uint8_p)
        if prob.G_csc is not None:
            Gs = prob.G_csc.c_struct()
            As = prob.A_csc.c_struct() if prob.p else None
            rc = h.lib.socp_b200_set_data_csc(h.ptr, _dp(prob.c), C.byref(As) if prob.p else None,
                                              _dp(prob.b) if prob.p else None, C.byref(Gs), _dp(prob.h), sing, flags)
            h.check(rc, "socp_b200_set_data_csc")
        else:
            rc = h.lib.socp_b200_set_data(h.ptr, _dp(prob.c), _dp(prob.A_cm) if prob.p else None,
                                          _dp(prob.b) if prob.p else None, _dp(prob.G_cm), _dp(prob.h), sing, flags)
            h.check(rc, "socp_b200_set_data")
        self._loaded_id = id(prob)

    def get_sing(self) -> np.ndarray:
        h = self.handle
        out = np.zeros(h.batch, dtype=np.uint8)
        h.check(h.lib.socp_b200_get_sing(h.ptr, out.ctypes.data_as(L.c_uint8_p)), "socp_b200_get_sing")
        return out

    def debug_fused_step(self, index: int, it: int, phase: int = 1) -> dict:
        """socp_b200_debug_fused_step: the inputs and outputs of one compute_scaling / setup_iter / solve_kkt
        (reference src/densesolver.jl:41-90) taken from inside the fused whole-solve kernel at Mehrotra iteration
        `it` of problem `index` (phase 1 = affine solve, 2 = combined solve).  H is returned as (n, n) with
        H[i, j] = (G'W^-2 G)[i, j]."""
        h = self.handle
        n, p, k = h.n, h.p, h.k
        out = dict(s=np.zeros(k), z=np.zeros(k), H=np.zeros((n, n)), dx=np.zeros(n), dy=np.zeros(max(p, 1)),
                   dz=np.zeros(k), ds=np.zeros(k), cx=np.zeros(n), cy=np.zeros(max(p, 1)), cz=np.zeros(k), cs=np.zeros(k))
        order = ("s", "z", "H", "dx", "dy", "dz", "ds", "cx", "cy", "cz", "cs")
        h.check(h.lib.socp_b200_debug_fused_step(h.ptr, int(index), int(it), int(phase), *[_dp(out[f]) for f in order]),
                "socp_b200_debug_fused_step")
        out["H"] = out["H"].T.copy()
        out["dy"], out["cy"] = out["dy"][:p], out["cy"][:p]
        return out

    def close(self):
        self.handle.close()


SolverState = BatchSolverState


def solve_socp_batch(prob: BatchProblem, ss: BatchSolverState, params: Optional[L.Params] = None,
                     reload: bool = True, want_iterates: bool = True) -> BatchResult:
    """solve_socp over the batch: upload (unless already resident), initial
    point + Mehrotra loop on the device, download."""
    h = ss.handle
    B, n, p, k = h.batch, h.n, h.p, h.k
    x = np.empty((B, n)) if want_iterates else None
    y = np.empty((B, p)) if want_iterates else None
    z = np.empty((B, k)) if want_iterates else None
    s = np.empty((B, k)) if want_iterates else None
    status = np.empty(B, dtype=np.int32)
    iters = np.empty(B, dtype=np.int32)
    pobj = np.empty(B)
    dobj = np.empty(B)
    prm = params if params is not None else default_params()
    if prob.G_csc is not None and (reload or ss._loaded_id != id(prob)):
        # the same from the reference's own storage (SparseMatrixCSC): only the stored values cross PCIe
        assert (prob.n, prob.p, prob.B) == (h.n, h.p, h.batch) and prob.cones == h.cones
        flags = (1 if prob.shared_A else 0) | (2 if prob.shared_G else 0)
        sing = None if prob.sing is None else prob.sing.ctypes.data_as(L.c_uint8_p)
        Gs = prob.G_csc.c_struct()
        As = prob.A_csc.c_struct() if p else None
        rc = h.lib.socp_b200_solve_host_csc(h.ptr, C.byref(prm), _dp(prob.c), C.byref(As) if p else None,
                                            _dp(prob.b) if p else None, C.byref(Gs), _dp(prob.h), sing, flags,
                                            _dp(x), _dp(y) if p else None, _dp(z), _dp(s),
                                            _ip(status), _ip(iters), _dp(pobj), _dp(dobj))
        h.check(rc, "socp_b200_solve_host_csc")
        ss._loaded_id = id(prob)
    elif prob.G_csc is None and (reload or ss._loaded_id != id(prob)):
        # Problem(...) + solve_socp(prob, ss) in one call: upload, solve and download overlap on the fused path
        assert (prob.n, prob.p, prob.B) == (h.n, h.p, h.batch) and prob.cones == h.cones
        flags = (1 if prob.shared_A else 0) | (2 if prob.shared_G else 0)
        sing = None if prob.sing is None else prob.sing.ctypes.data_as(L.c_uint8_p)
        rc = h.lib.socp_b200_solve_host(h.ptr, C.byref(prm), _dp(prob.c), _dp(prob.A_cm) if p else None,
                                        _dp(prob.b) if p else None, _dp(prob.G_cm), _dp(prob.h), sing, flags,
                                        _dp(x), _dp(y) if p else None, _dp(z), _dp(s),
                                        _ip(status), _ip(iters), _dp(pobj), _dp(dobj))
        h.check(rc, "socp_b200_solve_host")
        ss._loaded_id = id(prob)
    else:
        rc = h.lib.socp_b200_solve(h.ptr, C.byref(prm), _dp(x), _dp(y) if p else None, _dp(z), _dp(s),
                                   _ip(status), _ip(iters), _dp(pobj), _dp(dobj))
        h.check(rc, "socp_b200_solve")
    return BatchResult(x, y, z, s, status, iters, pobj, dobj, h.timings())


def solve_socp(prob: BatchProblem, ss: BatchSolverState, params: Optional[L.Params] = None) -> State:
    """solve_socp(prob, ss) -> State, reference src/solver.jl:40-152 (batch of one)."""
    assert prob.B == 1
    r = solve_socp_batch(prob, ss, params)
    return State(r.x[0], r.y[0], r.z[0], r.s[0], int(r.status[0]), int(r.iters[0]), float(r.pobj[0]), float(r.dobj[0]))


# ----------------------------------------------------------------------- step level
_cone_handles: dict = {}


def _cone_handle(cones, batch: int) -> _Handle:
    """Handle for the cone-only helpers (vprod, iprod, ...) which take no solver
    in the reference: layout with n = 1, p = 0."""
    key = (_as_cones(cones), int(batch))
    hd = _cone_handles.get(key)
    if hd is None:
        if len(_cone_handles) > 16:
            _cone_handles.pop(next(iter(_cone_handles))).close()
        hd = _Handle(1, 0, key[0], batch)
        _cone_handles[key] = hd
    return hd


def _batched(v, k: int) -> Tuple[np.ndarray, bool]:
    a = _f64(v)
    single = (a.ndim == 1)
    a = a.reshape(-1, k)
    return a, single


def _unbatch(a: np.ndarray, single: bool):
    return a[0] if single else a


def compute_scaling(cones, scaling: B200Scaling, s, z) -> B200Scaling:
    """compute_scaling(cones, scaling, s, z), reference src/scalings.jl:101-110."""
    h = scaling.handle
    sa, _ = _batched(s, h.k)
    za, _ = _batched(z, h.k)
    assert sa.shape == (h.batch, h.k) and za.shape == (h.batch, h.k)
    rc = h.lib.socp_b200_compute_scaling(h.ptr, _dp(sa), _dp(za), _dp(scaling.l), _dp(scaling.wbs), _dp(scaling.mu),
                                         _ip(scaling.fail))
    h.check(rc, "socp_b200_compute_scaling")
    return scaling


def setup_iter(solver: B200Solver, prob: BatchProblem, state, scaling: B200Scaling) -> np.ndarray:
    """setup_iter(solver, prob, state, scaling), reference src/densesolver.jl:41-52.
    Returns fail[B] (non-zero where cholesky! would throw)."""
    h = solver.handle
    fail = np.zeros(h.batch, dtype=np.int32)
    h.check(h.lib.socp_b200_setup_iter(h.ptr, _ip(fail)), "socp_b200_setup_iter")
    return fail


def solve_kkt(solver: B200Solver, prob: BatchProblem, state, scaling: B200Scaling, dx, dy, dz, ds, cx, cy, cz, cs):
    """solve_kkt(solver, prob, state, scaling, dx,dy,dz,ds, cx,cy,cz,cs), reference
    src/densesolver.jl:54-90: writes cx, cy, cz, cs in place."""
    h = solver.handle
    B, n, p, k = h.batch, h.n, h.p, h.k
    dxa, dza, dsa = _f64(dx).reshape(B, n), _f64(dz).reshape(B, k), _f64(ds).reshape(B, k)
    dya = _f64(dy).reshape(B, p)
    ox, oy, oz, os_ = np.empty((B, n)), np.empty((B, p)), np.empty((B, k)), np.empty((B, k))
    rc = h.lib.socp_b200_solve_kkt(h.ptr, _dp(dxa), _dp(dya) if p else None, _dp(dza), _dp(dsa),
                                   _dp(ox), _dp(oy) if p else None, _dp(oz), _dp(os_))
    h.check(rc, "socp_b200_solve_kkt")
    np.asarray(cx).reshape(B, n)[...] = ox
    if p:
        np.asarray(cy).reshape(B, p)[...] = oy
    np.asarray(cz).reshape(B, k)[...] = oz
    np.asarray(cs).reshape(B, k)[...] = os_


def _apply(fn_name: str, scaling: B200Scaling, inp, out):
    h = scaling.handle
    a, _ = _batched(inp, h.k)
    assert a.shape == (h.batch, h.k)
    o = np.empty_like(a)
    h.check(getattr(h.lib, fn_name)(h.ptr, _dp(a), _dp(o)), fn_name)
    np.asarray(out).reshape(h.batch, h.k)[...] = o
    return out


def scale_(cones, scaling: B200Scaling, inp, out):
    """scale!(cones, scl, s, op): op = W s, reference src/scalings.jl:159-165."""
    return _apply("socp_b200_scale", scaling, inp, out)


def iscale_(cones, scaling: B200Scaling, inp, out):
    """iscale!(cones, scl, s, op): op = W^-1 s, reference src/scalings.jl:167-173."""
    return _apply("socp_b200_iscale", scaling, inp, out)


def iwiw(cones, scaling: B200Scaling, inp, out):
    """out = iWiW * in (the dense gemv of reference src/densesolver.jl:86)."""
    return _apply("socp_b200_iwiw", scaling, inp, out)


def sqr_scaling(cones, scaling: B200Scaling):
    """The SqrScaling vectors of the scaling computed last (compute_scaling(cones, ::SqrScaling, s, z), reference
    src/sqrscalings.jl:177-185): returns (D, u, v), each (batch, k), with W^-2 = diag(D) + u u' - v v' cone by cone
    (the per-cone u_c, v_c of the reference have disjoint supports and are packed into one k-vector each)."""
    h = scaling.handle
    D, u, v = np.empty((h.batch, h.k)), np.empty((h.batch, h.k)), np.empty((h.batch, h.k))
    h.check(h.lib.socp_b200_sqr_scaling(h.ptr, _dp(D), _dp(u), _dp(v)), "socp_b200_sqr_scaling")
    return D, u, v


def make_e(cones, batch: int = 1):
    """make_e(cones), reference src/vectors.jl:7-38."""
    h = _cone_handle(cones, batch)
    o = np.empty((h.batch, h.k))
    h.check(h.lib.socp_b200_make_e(h.ptr, _dp(o)), "socp_b200_make_e")
    return o[0] if batch == 1 else o


def vprod(cones, u, v):
    """vprod(cones, u, v) = u o v, reference src/vectors.jl:54-81."""
    k = sum(c.dim for c in _as_cones(cones))
    ua, single = _batched(u, k)
    va, _ = _batched(v, k)
    h = _cone_handle(cones, ua.shape[0])
    o = np.empty_like(ua)
    h.check(h.lib.socp_b200_vprod(h.ptr, _dp(ua), _dp(va), _dp(o)), "socp_b200_vprod")
    return _unbatch(o, single)


def iprod(cones, lam, v):
    """iprod(cones, lam, v) = lam \\ v, reference src/vectors.jl:87-131."""
    k = sum(c.dim for c in _as_cones(cones))
    la, single = _batched(lam, k)
    va, _ = _batched(v, k)
    h = _cone_handle(cones, la.shape[0])
    o = np.empty_like(la)
    h.check(h.lib.socp_b200_iprod(h.ptr, _dp(la), _dp(va), _dp(o)), "socp_b200_iprod")
    return _unbatch(o, single)


def max_step(cones, x):
    """max_step(cones, x), reference src/mats.jl:1-28."""
    k = sum(c.dim for c in _as_cones(cones))
    xa, single = _batched(x, k)
    h = _cone_handle(cones, xa.shape[0])
    o = np.empty(xa.shape[0])
    h.check(h.lib.socp_b200_max_step(h.ptr, _dp(xa), _dp(o)), "socp_b200_max_step")
    return float(o[0]) if single else o


def compute_step(cones, l, ds, dz):
    """compute_step(cones, l, ds, dz), reference src/mats.jl:30-40."""
    k = sum(c.dim for c in _as_cones(cones))
    la, single = _batched(l, k)
    dsa, _ = _batched(ds, k)
    dza, _ = _batched(dz, k)
    h = _cone_handle(cones, la.shape[0])
    o = np.empty(la.shape[0])
    h.check(h.lib.socp_b200_compute_step(h.ptr, _dp(la), _dp(dsa), _dp(dza), _dp(o)), "socp_b200_compute_step")
    return float(o[0]) if single else o
