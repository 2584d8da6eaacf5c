"""CPU-side checks of the C-ABI library: it builds, loads, exports every symbol
include/socp_b200.h declares, and rejects bad layouts before touching CUDA.
No compute calls (no GPU here)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from socp_b200 import _lib as L
from socp_b200 import build as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "socp_b200.h")


@pytest.fixture(scope="module")
def lib():
    B.build()
    return L.load()


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(socp_b200_[a-z_A-Z0-9]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(L.SYMBOLS.keys())


def test_library_exports_every_declared_symbol(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", B.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\sT\s+(socp_b200_\w+)", out))
    for name in header_symbols():
        assert name in exported, name
        assert getattr(lib, name) is not None


def test_version_and_default_params(lib):
    assert lib.socp_b200_version() == 100
    p = L.Params()
    lib.socp_b200_default_params(C.byref(p))
    # the reference's literals: src/solver.jl:105, :122, :146, :91
    assert (p.max_iter, p.tol, p.step_damp, p.init_eps) == (40, 1e-5, 0.99, 1e-10)


def _layout(n, p, k, kinds, offs, dims):
    ka = np.array(kinds, dtype=np.int32)
    oa = np.array(offs, dtype=np.int32)
    da = np.array(dims, dtype=np.int32)
    lay = L.Layout(n, p, k, len(kinds), ka.ctypes.data_as(L.c_int32_p), oa.ctypes.data_as(L.c_int32_p),
                   da.ctypes.data_as(L.c_int32_p))
    return lay, (ka, oa, da)


@pytest.mark.parametrize("args", [
    (3, 0, 4, [0, 1], [0, 2], [1, 3]),      # offsets do not tile
    (3, 0, 5, [0, 1], [0, 1], [1, 3]),      # dims do not sum to k
    (3, 0, 4, [1, 0], [0, 3], [3, 1]),      # POC after SOC (reference src/scalings.jl:102)
    (0, 0, 4, [0, 1], [0, 1], [1, 3]),      # n = 0
    (3, 0, 4, [0, 7], [0, 1], [1, 3]),      # unknown kind
])
def test_create_rejects_bad_layout(lib, args):
    lay, keep = _layout(*args)
    h = L.H()
    rc = lib.socp_b200_create(C.byref(h), C.byref(lay), 1, None, 0)
    assert rc < 0 and not h.value
    assert lib.socp_b200_last_error(None)


def test_null_handle_is_an_error(lib):
    assert lib.socp_b200_destroy(None) < 0
    assert lib.socp_b200_solve_dev(None, None) < 0
    assert lib.socp_b200_setup_iter(None, None) < 0
