"""Build libheros_gpu.so in-tree with nvcc for sm_100a (see csrc/Makefile)."""
from __future__ import annotations

import os
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libheros_gpu.so")
CSRC = os.path.join(PKG_DIR, "csrc")


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(os.path.dirname(PKG_DIR), "include", "heros_gpu.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source of the package (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo)."""
    if force:
        subprocess.check_call(["make", "-C", CSRC, "clean"], stdout=subprocess.DEVNULL)
    if force or _stale():
        out = None if verbose else subprocess.DEVNULL
        subprocess.check_call(["make", "-C", CSRC, "-j4"], stdout=out, stderr=None if verbose else subprocess.STDOUT)
    return LIB_PATH
