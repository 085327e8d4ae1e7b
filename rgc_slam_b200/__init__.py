"""rgc_slam_b200 — B200-native scan-matching hot path of RGC-SLAM.

Host-side mirror of the reference's registration interface (``fast_gicp::FastGICP``,
rgc_slam/include/fast_gicp/gicp/fast_gicp.hpp:20-100) over the C-ABI in include/rgc_gicp.h.
The numeric work happens exclusively in librgc_gicp.so (hand-written sm_100a kernels); there is
no CPU fallback — constructing a context without the library or without a CUDA device raises.
"""
from .api import (FastGICP, FastVGICP, DIRECT1, DIRECT7, DIRECT27, ADDITIVE, ADDITIVE_WEIGHTED, MULTIPLICATIVE, Context, RgcError, knn, knn_self, voxel_grid, deskew, FeatureMap, lib, lib_path,  # noqa: F401
                  REG_NONE, REG_MIN_EIG, REG_NORMALIZED_MIN_EIG, REG_PLANE, REG_FROBENIUS,
                  OPT_GAUSS_NEWTON, OPT_LEVENBERG_MARQUARDT)
from . import synth  # noqa: F401

__all__ = ["FastGICP", "Context", "RgcError", "knn", "lib", "lib_path", "synth"]
