"""ctypes binding of include/socp_b200.h.  There is no fallback: if the shared
library is missing this raises, and every compute call needs a CUDA device."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_uint8_p = C.POINTER(C.c_uint8)
c_int64_p = C.POINTER(C.c_int64)


class Layout(C.Structure):
    _fields_ = [("n", C.c_int32), ("p", C.c_int32), ("k", C.c_int32), ("ncones", C.c_int32),
                ("cone_kind", c_int32_p), ("cone_offs", c_int32_p), ("cone_dim", c_int32_p)]


class Params(C.Structure):
    _fields_ = [("max_iter", C.c_int32), ("path", C.c_int32), ("tol", C.c_double),
                ("step_damp", C.c_double), ("init_eps", C.c_double)]


class Timings(C.Structure):
    _fields_ = [("h2d_ms", C.c_double), ("solve_ms", C.c_double), ("d2h_ms", C.c_double),
                ("kernel_launches", C.c_int64), ("iterations_max", C.c_int32), ("path_used", C.c_int32)]


class Csc(C.Structure):
    _fields_ = [("nnz", C.c_int64), ("colptr", c_int64_p), ("rowval", c_int64_p), ("nzval", c_double_p),
                ("index_base", C.c_int32)]


# every symbol include/socp_b200.h declares: name -> (restype, argtypes)
H = C.c_void_p
SYMBOLS = {
    "socp_b200_version": (C.c_int, []),
    "socp_b200_device_count": (C.c_int, []),
    "socp_b200_default_params": (None, [C.POINTER(Params)]),
    "socp_b200_create": (C.c_int, [C.POINTER(H), C.POINTER(Layout), C.c_int64, c_int32_p, C.c_int32]),
    "socp_b200_destroy": (C.c_int, [H]),
    "socp_b200_last_error": (C.c_char_p, [H]),
    "socp_b200_set_data": (C.c_int, [H, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_uint8_p, C.c_int32]),
    "socp_b200_set_data_csc": (C.c_int, [H, c_double_p, C.POINTER(Csc), c_double_p, C.POINTER(Csc), c_double_p,
                                         c_uint8_p, C.c_int32]),
    "socp_b200_solve": (C.c_int, [H, C.POINTER(Params), c_double_p, c_double_p, c_double_p, c_double_p,
                                  c_int32_p, c_int32_p, c_double_p, c_double_p]),
    "socp_b200_solve_host": (C.c_int, [H, C.POINTER(Params), c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                       c_uint8_p, C.c_int32, c_double_p, c_double_p, c_double_p, c_double_p,
                                       c_int32_p, c_int32_p, c_double_p, c_double_p]),
    "socp_b200_solve_host_csc": (C.c_int, [H, C.POINTER(Params), c_double_p, C.POINTER(Csc), c_double_p, C.POINTER(Csc),
                                           c_double_p, c_uint8_p, C.c_int32, c_double_p, c_double_p, c_double_p,
                                           c_double_p, c_int32_p, c_int32_p, c_double_p, c_double_p]),
    "socp_b200_solve_dev": (C.c_int, [H, C.POINTER(Params)]),
    "socp_b200_get_results": (C.c_int, [H, c_double_p, c_double_p, c_double_p, c_double_p,
                                        c_int32_p, c_int32_p, c_double_p, c_double_p]),
    "socp_b200_get_sing": (C.c_int, [H, c_uint8_p]),
    "socp_b200_timings": (C.c_int, [H, C.POINTER(Timings)]),
    "socp_b200_compute_scaling": (C.c_int, [H, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_int32_p]),
    "socp_b200_setup_iter": (C.c_int, [H, c_int32_p]),
    "socp_b200_solve_kkt": (C.c_int, [H] + [c_double_p] * 8),
    "socp_b200_scale": (C.c_int, [H, c_double_p, c_double_p]),
    "socp_b200_iscale": (C.c_int, [H, c_double_p, c_double_p]),
    "socp_b200_iwiw": (C.c_int, [H, c_double_p, c_double_p]),
    "socp_b200_sqr_scaling": (C.c_int, [H, c_double_p, c_double_p, c_double_p]),
    "socp_b200_make_e": (C.c_int, [H, c_double_p]),
    "socp_b200_vprod": (C.c_int, [H, c_double_p, c_double_p, c_double_p]),
    "socp_b200_iprod": (C.c_int, [H, c_double_p, c_double_p, c_double_p]),
    "socp_b200_max_step": (C.c_int, [H, c_double_p, c_double_p]),
    "socp_b200_compute_step": (C.c_int, [H, c_double_p, c_double_p, c_double_p, c_double_p]),
    "socp_b200_get_H": (C.c_int, [H, c_double_p]),
    "socp_b200_get_L": (C.c_int, [H, c_double_p]),
    "socp_b200_debug_fused_step": (C.c_int, [H, C.c_int64, C.c_int32, C.c_int32] + [c_double_p] * 11),
    "socp_b200_profile_step": (C.c_int, [H, C.c_int32, C.c_int32, c_double_p]),
}

_lib = None


def lib_path() -> str:
    # SOCP_B200_LIB lets tools/ load the profiling build; the product default is the in-tree library
    return os.environ.get("SOCP_B200_LIB", _build.LIB_PATH)


def load():
    """Load libsocp_b200.so (built in-tree by build.py / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
