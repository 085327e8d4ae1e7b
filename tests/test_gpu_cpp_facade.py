"""GPU: the header-only C++ facade (include/rgc/fast_gicp.hpp) driven exactly like the reference
call site (RGC_odometer.cpp:998-1015), compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

from conftest import rot_angle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_facade_matches_oracle(small_pair, tmp_path):
    from oracle import oracle as orc
    src, tgt, _ = small_pair
    exe = str(tmp_path / "facade_smoke")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    libdir = os.path.join(ROOT, "rgc_slam_b200")
    subprocess.check_call([cxx, "-std=c++17", "-O2", "-o", exe, os.path.join(ROOT, "tests", "cpp", "facade_smoke.cpp"),
                           "-L", libdir, "-lrgc_gicp", f"-Wl,-rpath,{libdir}"])
    tgt.tofile(tmp_path / "tgt.bin")
    src.tofile(tmp_path / "src.bin")
    out = subprocess.run([exe, str(tmp_path / "tgt.bin"), str(tmp_path / "src.bin")], capture_output=True, text=True, check=True).stdout.split("\n")
    head = out[0].split()
    T = np.array([[float(v) for v in out[1 + r].split()] for r in range(4)])
    o = orc.FastGICP(max_iterations=25, corr_dist=2.0, transformation_epsilon=1e-6)
    o.setInputTarget(tgt)
    o.setInputSource(src)
    To = o.align()
    assert int(head[1]) == int(o.last["converged"]) and int(head[3]) == o.last["iterations"]
    assert np.abs(T[:3, 3] - To[:3, 3]).max() < 1e-4 and rot_angle(T[:3, :3], To[:3, :3]) < 1e-5
    assert abs(float(head[5]) - o.getFitnessScore()) < 1e-6 * o.getFitnessScore()
    assert int(head[7]) == len(src) and int(head[9]) == 1
    # second block: rgc::FastVGICP driven exactly like RGC_odometer.cpp:998-1011
    vhead = out[5].split()
    Tv = np.array([[float(v) for v in out[6 + r].split()] for r in range(4)])
    ov = orc.FastVGICP(resolution=1.0, max_iterations=25, transformation_epsilon=1e-6)
    ov.setInputTarget(tgt)
    ov.setInputSource(src)
    Tov = ov.align()
    assert vhead[0] == "vgicp" and int(vhead[2]) == int(ov.last["converged"]) and int(vhead[4]) == ov.last["iterations"]
    assert np.abs(Tv[:3, 3] - Tov[:3, 3]).max() < 1e-4 and rot_angle(Tv[:3, :3], Tov[:3, :3]) < 1e-5
    # third block: the fused front end (voxel filters + setInput*) equals the Python binding's
    fhead = out[10].split()
    Tf = np.array([[float(v) for v in out[11 + r].split()] for r in range(4)])
    import rgc_slam_b200 as rgc
    S, T_ = src.copy(), tgt.copy()
    S[:, 3] = np.arange(len(S), dtype=np.float32)      # facade_smoke.cpp's load() stores the point index as intensity
    T_[:, 3] = np.arange(len(T_), dtype=np.float32)
    g = rgc.FastGICP()
    g.setMaxCorrespondenceDistance(2.0)
    nt = g.setInputTargetFiltered(T_, 0.3)
    ns = g.setInputSourceFiltered(S, 0.2)
    Tg = g.align()
    assert fhead[0] == "filtered" and int(fhead[6]) == ns and int(fhead[8]) == nt and int(fhead[10]) == ns
    assert int(fhead[4]) == g.last_result["iterations"]
    assert np.abs(Tf - Tg).max() < 1e-6
    # fourth block: swapSourceAndTarget swaps the host-side cloud handles too
    shead = out[15].split()
    assert shead[0] == "swap" and int(shead[4]) == len(tgt) and int(shead[6]) == len(tgt) and int(shead[8]) == 1
