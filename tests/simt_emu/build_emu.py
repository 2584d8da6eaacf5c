"""Builds tests/_build/libsocp_emu.so: the whole-solve kernel of socp.jl_b200/csrc/fused_v3.cuh compiled for the HOST
against the SIMT emulator (tests/simt_emu/simt_emu.h).  TEST INFRASTRUCTURE ONLY -- lets `pytest -m "not gpu"` run the
kernel's logic against the oracle without a GPU.  Nothing under socp.jl_b200/ uses it."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "socp.jl_b200", "csrc")
OUT_DIR = os.path.join(ROOT, "tests", "_build")
OUT = os.path.join(OUT_DIR, "libsocp_emu.so")
SRCS = [os.path.join(HERE, f) for f in ("emu_fused3.cpp", "emu_fused_lane.cpp")]
DEPS = SRCS + [os.path.join(HERE, "simt_emu.h")] + \
       [os.path.join(CSRC, f) for f in ("common.cuh", "fused_common.cuh", "fused_v3.cuh", "fused_lane.cuh", "fused_lane_dev.cuh")]


def build(force: bool = False) -> str:
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in DEPS):
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["g++", "-O2", "-g", "-std=c++17", "-DSOCP_SIMT_EMU", "-I" + HERE, "-I" + CSRC, "-shared", "-fPIC",
           "-o", OUT] + SRCS
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building the emulator library")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
