"""ctypes wrapper of the C oracle (oracle/socp_oracle.c).  TEST INFRASTRUCTURE ONLY
(tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs)."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from . import build_oracle

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class _Layout(C.Structure):
    _fields_ = [("n", C.c_int), ("p", C.c_int), ("k", C.c_int), ("ncones", C.c_int),
                ("kind", C.POINTER(C.c_int)), ("offs", C.POINTER(C.c_int)), ("dim", C.POINTER(C.c_int))]


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = build_oracle.build()
        _lib = C.CDLL(path)
        _lib.oc_solve_batch.restype = C.c_int
        _lib.oc_kkt_step.restype = C.c_int
        _lib.oc_num_threads.restype = C.c_int
    return _lib


def _layout(n, p, cones):
    kind = np.array([c[0] for c in cones], dtype=np.int32)
    offs = np.array([c[1] for c in cones], dtype=np.int32)
    dim = np.array([c[2] for c in cones], dtype=np.int32)
    k = int(dim.sum())
    lay = _Layout(n, p, k, len(cones), kind.ctypes.data_as(C.POINTER(C.c_int)),
                  offs.ctypes.data_as(C.POINTER(C.c_int)), dim.ctypes.data_as(C.POINTER(C.c_int)))
    return lay, (kind, offs, dim), k


def _d(a):
    return a.ctypes.data_as(_dp)


def solve_batch(c, A_cm, b, G_cm, h, cones, sing=None, max_iter=40, tol=1e-5, step_damp=0.99, init_eps=1e-10,
                nthreads: int = 0, shared_A=False, shared_G=False):
    """c (B,n); A_cm (B,n,p) = column-major p x n per problem (or (n,p) shared); b (B,p);
    G_cm (B,n,k) (or (n,k) shared); h (B,k).  Returns dict of arrays."""
    c = np.ascontiguousarray(c, dtype=np.float64)
    B, n = c.shape
    b = np.ascontiguousarray(b, dtype=np.float64).reshape(B, -1)
    p = b.shape[1]
    lay, keep, k = _layout(n, p, cones)
    G_cm = np.ascontiguousarray(G_cm, dtype=np.float64)
    A_cm = np.ascontiguousarray(A_cm, dtype=np.float64) if p else np.zeros(1)
    h = np.ascontiguousarray(h, dtype=np.float64)
    x, y, z, s = np.zeros((B, n)), np.zeros((B, max(p, 1))), np.zeros((B, k)), np.zeros((B, k))
    y = np.zeros((B, p)) if p else np.zeros((B, 0))
    ybuf = y if p else np.zeros(1)
    status, iters = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
    pobj, dobj = np.zeros(B), np.zeros(B)
    sg = None
    if sing is not None:
        sg = np.ascontiguousarray(sing, dtype=np.uint8)
    lib().oc_solve_batch(C.byref(lay), C.c_int64(B), _d(c), _d(A_cm), C.c_int64(0 if shared_A else p * n), _d(b) if p else _d(np.zeros(1)),
                         _d(G_cm), C.c_int64(0 if shared_G else k * n), _d(h),
                         sg.ctypes.data_as(C.POINTER(C.c_uint8)) if sg is not None else None,
                         C.c_int(max_iter), C.c_double(tol), C.c_double(step_damp), C.c_double(init_eps), C.c_int(nthreads),
                         _d(x), _d(ybuf), _d(z), _d(s), status.ctypes.data_as(_ip), iters.ctypes.data_as(_ip), _d(pobj), _d(dobj))
    return dict(x=x, y=y, z=z, s=s, status=status, iters=iters, pobj=pobj, dobj=dobj)


def kkt_step(A, G, cones, sing, s, z, dx, dy, dz, ds):
    """One compute_scaling + setup_iter + solve_kkt for a single problem; A (p,n), G (k,n) logical."""
    G = np.asarray(G, dtype=np.float64)
    k, n = G.shape
    A = np.asarray(A, dtype=np.float64).reshape(-1, n)
    p = A.shape[0]
    lay, keep, k2 = _layout(n, p, cones)
    assert k2 == k
    Gc = np.ascontiguousarray(G.T)
    Ac = np.ascontiguousarray(A.T) if p else np.zeros(1)
    f = lambda v: np.ascontiguousarray(v, dtype=np.float64)
    s, z, dx, dz, ds = f(s), f(z), f(dx), f(dz), f(ds)
    dy = f(dy) if p else np.zeros(1)
    cx, cy, cz, cs = np.zeros(n), np.zeros(max(p, 1)), np.zeros(k), np.zeros(k)
    lam, wbs, mu = np.zeros(k), np.zeros(k), np.zeros(len(cones))
    rc = lib().oc_kkt_step(C.byref(lay), _d(Ac), _d(Gc), C.c_int(1 if sing else 0), _d(s), _d(z), _d(dx), _d(dy), _d(dz), _d(ds),
                           _d(cx), _d(cy), _d(cz), _d(cs), _d(lam), _d(wbs), _d(mu))
    return dict(rc=rc, cx=cx, cy=cy[:p], cz=cz, cs=cs, l=lam, wbs=wbs, mu=mu)


def num_threads() -> int:
    return lib().oc_num_threads()
