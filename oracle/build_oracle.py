"""Builds the C oracle (oracle/socp_oracle.c) into oracle/_build/libsocp_oracle.so.
TEST INFRASTRUCTURE: the checker and the timed CPU baseline, never the product."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "socp_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libsocp_oracle.so")


def build(force: bool = False) -> str:
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    # -march=x86-64-v3 (AVX2+FMA) rather than native: the .so is built here and run on the GPU box
    cmd = ["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-fPIC", "-shared", "-o", OUT, SRC, "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("gcc failed building the C oracle")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
