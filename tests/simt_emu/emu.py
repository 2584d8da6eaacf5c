"""ctypes wrapper of tests/_build/libsocp_emu.so (the fused kernel on the SIMT emulator).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import build_emu  # noqa: E402

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_emu.build())
        _lib.emu_fused3_solve.restype = C.c_int
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def plan(n, p, cones, rowcol):
    kind = np.array([c[0] for c in cones], dtype=np.int32)
    offs = np.array([c[1] for c in cones], dtype=np.int32)
    dim = np.array([c[2] for c in cones], dtype=np.int32)
    rc = np.ascontiguousarray(rowcol, dtype=np.int32)
    out = np.zeros(8, dtype=np.int32)
    lib().emu_fused3_plan(n, p, int(dim.sum()), len(cones), _i(kind), _i(offs), _i(dim), _i(rc), _i(out))
    return dict(zip(("fits", "smem", "ctas_per_sm", "d0", "kd", "nsing", "ident", "nb"), out.tolist()))


def solve(c, A_cm, b, G_cm, h, cones, sing=None, rowcol=None, generic=False, sing_detect=False, verify=False, order=0,
          max_iter=40, tol=1e-5, step_damp=0.99, init_eps=1e-10, dbg=None, grid_cap=4, shared_G=False, shared_A=False,
          teams4=False, align=False):
    """Same conventions as oracle.c_oracle.solve_batch.  dbg = (problem, iteration[, phase]) -> also returns the debug dump."""
    c = np.ascontiguousarray(c, dtype=np.float64)
    B, n = c.shape
    b = np.ascontiguousarray(b, dtype=np.float64).reshape(B, -1)
    p = b.shape[1]
    kind = np.array([cn[0] for cn in cones], dtype=np.int32)
    offs = np.array([cn[1] for cn in cones], dtype=np.int32)
    dim = np.array([cn[2] for cn in cones], dtype=np.int32)
    k = int(dim.sum())
    G_cm = np.ascontiguousarray(G_cm, dtype=np.float64)
    A_cm = np.ascontiguousarray(A_cm, dtype=np.float64) if p else np.zeros(1)
    h = np.ascontiguousarray(h, dtype=np.float64)
    x, y, z, s = np.zeros((B, n)), np.zeros((B, max(p, 1))), np.zeros((B, k)), np.zeros((B, k))
    status, iters = np.full(B, -99, dtype=np.int32), np.zeros(B, dtype=np.int32)
    pobj, dobj = np.zeros(B), np.zeros(B)
    sing_out = np.zeros(B, dtype=np.uint8)
    npat = np.zeros(1, dtype=np.int32)
    sg = np.ascontiguousarray(sing, dtype=np.uint8) if sing is not None else None
    rc = np.ascontiguousarray(rowcol, dtype=np.int32) if rowcol is not None else None
    dbuf = np.zeros(2 * k + n * n + 2 * (n + p + 2 * k)) if dbg is not None else None
    flags = (1 if generic else 0) | (2 if sing_detect else 0) | (4 if verify else 0) | (8 if teams4 else 0) | (16 if align else 0)
    r = lib().emu_fused3_solve(n, p, k, len(cones), _i(kind), _i(offs), _i(dim), B, _d(c), _d(A_cm),
                               C.c_int64(0 if shared_A else p * n), _d(b) if p else _d(np.zeros(1)), _d(G_cm),
                               C.c_int64(0 if shared_G else k * n), _d(h),
                               sg.ctypes.data_as(C.POINTER(C.c_uint8)) if sg is not None else None,
                               _i(rc) if rc is not None else None, flags, order, max_iter, C.c_double(tol),
                               C.c_double(step_damp), C.c_double(init_eps), _d(x), _d(y), _d(z), _d(s), _i(status),
                               _i(iters), _d(pobj), _d(dobj), sing_out.ctypes.data_as(C.POINTER(C.c_uint8)), _i(npat),
                               _d(dbuf) if dbuf is not None else None, dbg[0] if dbg else -1, dbg[1] if dbg else -1,
                               (dbg[2] if len(dbg) > 2 else 1) if dbg else 1, grid_cap)
    if r != 0:
        raise RuntimeError("layout does not fit the fused kernel (emulated plan)")
    out = dict(x=x, y=y[:, :p], z=z, s=s, status=status, iters=iters, pobj=pobj, dobj=dobj, sing=sing_out,
               npattern=int(npat[0]))
    if dbuf is not None:
        o = 0
        def take(m):
            nonlocal o
            v = dbuf[o:o + m]
            o += m
            return v
        out["dbg"] = dict(s=take(k), z=take(k), H=take(n * n).reshape(n, n).T, dx=take(n), dy=take(p), dz=take(k),
                          ds=take(k), cx=take(n), cy=take(p), cz=take(k), cs=take(k))
    return out


def solve_lane(c, G_cm, h, cones, lpw=8, order=0, grid_cap=0, max_iter=40, tol=1e-5, step_damp=0.99, init_eps=1e-10,
               shared_G=False):
    """The lane-per-problem kernel k_fused_lane (socp.jl_b200/csrc/fused_lane.cuh) on the emulator; p = 0 layouts."""
    c = np.ascontiguousarray(c, dtype=np.float64)
    B, n = c.shape
    kind = np.array([cn[0] for cn in cones], dtype=np.int32)
    offs = np.array([cn[1] for cn in cones], dtype=np.int32)
    dim = np.array([cn[2] for cn in cones], dtype=np.int32)
    k = int(dim.sum())
    G_cm = np.ascontiguousarray(G_cm, dtype=np.float64)
    h = np.ascontiguousarray(h, dtype=np.float64)
    x, z, s = np.zeros((B, n)), np.zeros((B, k)), np.zeros((B, k))
    status, iters = np.full(B, -99, dtype=np.int32), np.zeros(B, dtype=np.int32)
    pobj, dobj = np.zeros(B), np.zeros(B)
    f = lib().emu_fused_lane_solve
    f.restype = C.c_int
    r = f(n, k, len(cones), _i(kind), _i(offs), _i(dim), B, _d(c), _d(G_cm), C.c_int64(0 if shared_G else k * n), _d(h),
          lpw, order, grid_cap, max_iter, C.c_double(tol), C.c_double(step_damp), C.c_double(init_eps), _d(x), _d(z),
          _d(s), _i(status), _i(iters), _d(pobj), _d(dobj))
    if r != 0:
        raise RuntimeError("no lane-per-problem instantiation takes this layout")
    return dict(x=x, y=np.zeros((B, 0)), z=z, s=s, status=status, iters=iters, pobj=pobj, dobj=dobj)


def plan_lane(n, p, cones):
    """fl_plan (host side of the lane-per-problem kernel) for a layout."""
    kind = np.array([c[0] for c in cones], dtype=np.int32)
    offs = np.array([c[1] for c in cones], dtype=np.int32)
    dim = np.array([c[2] for c in cones], dtype=np.int32)
    out = np.zeros(24, dtype=np.int32)
    lib().emu_fused_lane_plan(n, p, int(dim.sum()), len(cones), _i(kind), _i(offs), _i(dim), _i(out))
    ng = int(out[5])
    return dict(fits=bool(out[0]), shape=int(out[1]), pps=int(out[2]), smem=int(out[3]), rs=int(out[4]),
                groups=[(int(out[6 + 2 * i]), int(out[7 + 2 * i])) for i in range(ng)])
