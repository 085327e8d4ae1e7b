"""CPU: the C-ABI library loads and exports every symbol include/rgc_gicp.h declares; without a
GPU the product fails loudly instead of falling back."""
import ctypes
import os
import re

import pytest

import rgc_slam_b200 as rgc
from rgc_slam_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    out = []
    for h in ("rgc_gicp.h", "rgc_features.h", "rgc_preprocess.h", "rgc_mapping.h", "rgc_batch.h"):
        p = os.path.join(ROOT, "include", h)
        if os.path.exists(p):
            text = re.sub(r"/\*.*?\*/", "", open(p).read(), flags=re.S)
            out += re.findall(r"\b(rgc_[a-z0-9_]+)\s*\(", text)
    return sorted(set(out))


def test_library_exports_every_declared_symbol():
    from rgc_slam_b200 import build
    path = build.build()  # nvcc cross-compiles sm_100a without a GPU
    L = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ but not exported by librgc_gicp.so"
    for n in api.EXPORTED_SYMBOLS:
        assert n in names


def test_only_sm100a_code_is_embedded():
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "-lelf", api.lib_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(rgc.RgcError):
        rgc.Context(0)
    with pytest.raises(rgc.RgcError):
        rgc.FastGICP()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "rgc_slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for pat in (r"^\s*(from|import)\s+oracle", r"#include\s+[\"<][^\">]*orc_", r"liboracle", r"orc_[a-z]+\("):
                    assert not re.search(pat, text, flags=re.M), f"{f} uses the oracle ({pat})"
