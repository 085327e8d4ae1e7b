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


def test_headers_are_plain_c_and_the_integration_example_links(tmp_path):
    """The boundary is a C ABI: every header under include/ must compile as C99 (no C++ in the signatures) and a C
    translation unit that spells out the call sequences of INTEGRATION.md must link against librgc_gicp.so (nothing is
    executed: there is no GPU here)."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("gcc not on PATH")
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stddef.h>
#include <stdint.h>
#include <float.h>
#include "rgc_gicp.h"
#include "rgc_batch.h"
#include "rgc_features.h"
#include "rgc_preprocess.h"
#include "rgc_mapping.h"
int odometry(const float* tgt, size_t n_t, const float* src, size_t n_s, const float* guess, float* T) {
  rgc_ctx* ctx; rgc_reg* reg; rgc_params prm; rgc_result res; double fit;
  if (rgc_ctx_create(0, &ctx) != RGC_OK) return 1;
  rgc_reg_create(ctx, &reg);
  rgc_params_default(&prm); prm.max_iterations = 25; prm.max_correspondence_distance = 2.f;
  rgc_reg_set_params(reg, &prm);
  rgc_reg_set_target(reg, tgt, n_t, 32, 1); rgc_reg_set_source(reg, src, n_s, 32, 2);
  rgc_reg_align(reg, guess, T, &res, NULL); rgc_reg_fitness(reg, DBL_MAX, &fit);
  rgc_reg_destroy(reg); rgc_ctx_destroy(ctx);
  return res.converged ? 0 : 2;
}
int loop_closures(rgc_ctx* ctx, const rgc_pair* pairs, size_t n, rgc_pair_result* out) {
  rgc_params prm; rgc_params_default(&prm);
  return rgc_batch_align(ctx, &prm, pairs, n, 1, DBL_MAX, 0, out);
}
int sharded(rgc_ctx* ctx, rgc_reg* reg, int rank, int world, char* id, const float* map, size_t n) {
  rgc_comm* comm; size_t n_local;
  if (rank == 0) rgc_comm_unique_id(id);
  if (rgc_comm_create(ctx, id, rank, world, &comm) != RGC_OK) return 1;
  rgc_reg_set_comm(reg, comm);
  rgc_reg_set_target_slab(reg, map, n, 16, 0, -10.f, 10.f, 4.f, 7, &n_local, NULL);
  return rgc_comm_transport(comm);
}
int main(void) { return 0; }
''')
    inc = os.path.join(ROOT, "include")
    libdir = os.path.dirname(api.lib_path())
    exe = tmp_path / "abi"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", f"-I{inc}", str(src), "-o", str(exe), f"-L{libdir}", "-lrgc_gicp",
                        f"-Wl,-rpath,{libdir}", "-Wl,--unresolved-symbols=ignore-in-shared-libs"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
