"""Builds libsocp_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)                     # socp.jl_b200/
CSRC = os.path.join(ROOT, "csrc")
LIB_DIR = os.path.join(ROOT, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libsocp_b200.so")
PROF_LIB_PATH = os.path.join(LIB_DIR, "libsocp_b200_prof.so")   # -DSOCP_PHASE_TIMING build (tools/ only)

NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC"]
# translation units (compiled in parallel, linked into one shared library) and the headers each one includes
UNITS = {
    "solver.cu": None,                                                 # everything
    "fused2.cu": ["common.cuh", "fused_common.cuh", "fused_v2.cuh"],
    "fused3.cu": ["common.cuh", "fused_common.cuh", "fused_v3.cuh"],
    "fused3_dyn.cu": ["common.cuh", "fused_common.cuh", "fused_v3.cuh"],
    "fused_lane.cu": ["common.cuh", "fused_common.cuh", "fused_lane.cuh", "fused_lane_dev.cuh"],
    "lane_jit.cu": ["common.cuh", "fused_common.cuh", "fused_lane.cuh", "fused_lane_dev.cuh", "fused_v3.cuh", "fused_v2.cuh"],
    "syrk_tma.cu": ["common.cuh", "syrk_tma.cuh"],
}


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + [os.path.join(os.path.dirname(ROOT), "include", "socp_b200.h")]
    return any(os.path.getmtime(s) > t for s in deps)


def _compile(out_path: str, extra, verbose: bool = False, force: bool = False) -> None:
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    obj_dir = os.path.join(LIB_DIR, "obj" + ("_prof" if extra else ""))
    os.makedirs(obj_dir, exist_ok=True)
    procs = []
    objs = []
    header = os.path.join(os.path.dirname(ROOT), "include", "socp_b200.h")
    for u, deps in UNITS.items():
        obj = os.path.join(obj_dir, u.replace(".cu", ".o"))
        objs.append(obj)
        dep_paths = [os.path.join(CSRC, u)] + ([os.path.join(CSRC, d) for d in deps] if deps is not None else sources() + [header])
        if not force and os.path.exists(obj) and all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in dep_paths):
            continue                                                   # this unit is up to date
        cmd = [nvcc] + NVCC_FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, os.path.join(CSRC, u)]
        procs.append((u, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    for u, obj, pr in procs:
        out, err = pr.communicate()
        if pr.returncode != 0:
            sys.stderr.write(out + err)
            raise RuntimeError("nvcc failed compiling " + u)
        if verbose:
            sys.stderr.write(err)
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out_path] + objs + ["-ldl"],
                         capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking " + os.path.basename(out_path))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    _compile(LIB_PATH, [], verbose, force)
    return LIB_PATH


def build_prof() -> str:
    """Profiling variant with per-phase clock64() counters in the fused kernels (tools/phase_timing.py)."""
    _compile(PROF_LIB_PATH, ["-DSOCP_PHASE_TIMING"])
    return PROF_LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
