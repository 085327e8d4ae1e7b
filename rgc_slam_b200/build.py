"""Build librgc_gicp.so (hand-written sm_100a kernels + C-ABI) in-tree with nvcc.

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU box
with the gpurun snapshot.  `python -m rgc_slam_b200.build` rebuilds unconditionally.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "librgc_gicp.so")
SOURCES = ["rgc_gicp.cu", "rgc_features.cu"]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def sources() -> list[str]:
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(PKG, "..", "include", "rgc_gicp.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS
    if os.path.exists("/usr/bin/g++"):
        cmd += ["-ccbin", "/usr/bin/g++"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += os.environ.get("RGC_NVCC_EXTRA", "").split()  # tuning experiments: -DRGC_KT_PEND=8 ...
    out = os.environ.get("RGC_LIB_OUT", LIB)  # tuning experiments: a variant build beside the product library
    cmd += ["-o", out] + sources()
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
