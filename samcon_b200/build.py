"""Build libsamcon_b200.so in-tree with nvcc for sm_100a (cross-compiles on a box without a GPU).

    python -m samcon_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import subprocess
import sys

_CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(_CSRC, "libsamcon_b200.so")
SOURCES = ["samcon_b200.cu"]
HEADERS = ["sc_core.cuh", "sc_tables.h", os.path.join("..", "..", "include", "samcon_b200.h")]
# Not --use_fast_math (it would replace sinf/cosf/atan2f by the low-accuracy intrinsics on the trajectory), only its
# division / square-root / denormal parts: 2-ulp reciprocal and rsqrt instead of the IEEE sequences (3.5 % on the walk
# window, 7 % without self collision; parity tolerances are 1e-3).
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-prec-div=false",
              "-prec-sqrt=false", "-ftz=true", "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(_CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    # (on a box without nvcc -- never the case in this image -- a prebuilt library is used as it is)
    nvcc_path = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    if (not force and not _stale()) or (not os.path.exists(nvcc_path) and os.path.exists(LIB)):
        return LIB
    nvcc = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc, *NVCC_FLAGS, "-o", LIB, *[os.path.join(_CSRC, s) for s in SOURCES]]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed")
    with open(os.path.join(_CSRC, "ptxas_info.txt"), "w") as f:
        f.write(r.stderr)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv)
    print(LIB)
