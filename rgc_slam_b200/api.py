"""ctypes binding of include/rgc_gicp.h and a Python mirror of fast_gicp::FastGICP.

Method names, argument meaning and defaults follow the reference class
(rgc_slam/include/fast_gicp/gicp/fast_gicp.hpp:51-76, lsq_registration.hpp:51-61) and the
pcl::Registration calls its users make (rgc_slam/src/RGC_odometer.cpp:998-1011), so a parity
test reads like the reference call site.  Matrices cross this layer as row-major numpy arrays
(numpy's natural layout); the C-ABI itself is column-major (Eigen's) — transposed here.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref

import numpy as np

REG_NONE, REG_MIN_EIG, REG_NORMALIZED_MIN_EIG, REG_PLANE, REG_FROBENIUS = range(5)
OPT_GAUSS_NEWTON, OPT_LEVENBERG_MARQUARDT = 0, 1

_PKG = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class RgcError(RuntimeError):
    pass


class _Params(C.Structure):
    _fields_ = [("max_iterations", C.c_int), ("rotation_epsilon", C.c_double), ("transformation_epsilon", C.c_double),
                ("max_correspondence_distance", C.c_float), ("k_correspondences", C.c_int), ("regularization", C.c_int),
                ("optimizer", C.c_int), ("lm_max_iterations", C.c_int), ("lm_init_lambda_factor", C.c_double),
                ("lm_debug_print", C.c_int), ("grid_cell", C.c_float)]


class _Result(C.Structure):
    _fields_ = [("converged", C.c_int), ("iterations", C.c_int), ("n_linearize", C.c_int), ("n_compute_error", C.c_int),
                ("n_inliers", C.c_int), ("final_error", C.c_double), ("final_hessian", C.c_double * 36), ("device_ms", C.c_float)]


def lib_path() -> str:
    # RGC_LIB: A/B timing of two builds of this same library inside one GPU session (tools/)
    return os.environ.get("RGC_LIB") or os.path.join(_PKG, "librgc_gicp.so")


def lib():
    """Load librgc_gicp.so.  Fails loudly if the CUDA extension has not been built."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise RgcError(f"{path} is missing: build it with `python -m rgc_slam_b200.build` "
                           "(nvcc, sm_100a).  There is no CPU fallback.")
        L = C.CDLL(path)
        vp, sz, u64 = C.c_void_p, C.c_size_t, C.c_uint64
        L.rgc_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.rgc_ctx_destroy.argtypes = [vp]
        L.rgc_last_error.argtypes = [vp]
        L.rgc_last_error.restype = C.c_char_p
        L.rgc_ctx_synchronize.argtypes = [vp]
        L.rgc_ctx_stream.argtypes = [vp]
        L.rgc_ctx_stream.restype = vp
        L.rgc_ctx_set_profiling.argtypes = [vp, C.c_int]
        L.rgc_ctx_last_kernel_ms.argtypes = [vp, vp]
        L.rgc_ctx_launch_count.argtypes = [vp]
        L.rgc_ctx_launch_count.restype = u64
        L.rgc_reg_create.argtypes = [vp, C.POINTER(vp)]
        L.rgc_reg_destroy.argtypes = [vp]
        L.rgc_params_default.argtypes = [C.POINTER(_Params)]
        L.rgc_reg_set_params.argtypes = [vp, C.POINTER(_Params)]
        L.rgc_reg_get_params.argtypes = [vp, C.POINTER(_Params)]
        for name in ("rgc_reg_set_source", "rgc_reg_set_target", "rgc_reg_set_source_device", "rgc_reg_set_target_device"):
            getattr(L, name).argtypes = [vp, vp, sz, sz, u64]
        for name in ("rgc_reg_swap_source_and_target", "rgc_reg_clear_source", "rgc_reg_clear_target", "rgc_reg_sync_inputs"):
            getattr(L, name).argtypes = [vp]
        for name in ("rgc_reg_set_source_covs", "rgc_reg_set_target_covs", "rgc_reg_get_source_covs", "rgc_reg_get_target_covs"):
            getattr(L, name).argtypes = [vp, vp, sz]
        L.rgc_reg_align.argtypes = [vp, vp, vp, C.POINTER(_Result), vp]
        L.rgc_reg_linearize.argtypes = [vp, vp, C.POINTER(C.c_double), vp, vp]
        L.rgc_reg_compute_error.argtypes = [vp, vp, C.POINTER(C.c_double)]
        L.rgc_reg_get_correspondences.argtypes = [vp, vp, vp]
        L.rgc_reg_fitness.argtypes = [vp, C.c_double, C.POINTER(C.c_double)]
        L.rgc_reg_get_final_transformation.argtypes = [vp, vp]
        L.rgc_knn.argtypes = [vp, vp, sz, sz, vp, sz, sz, C.c_int, vp, vp, C.c_float]
        L.rgc_reg_stage_ms.argtypes = [vp, vp]
        L.rgc_knn_self.argtypes = [vp, vp, sz, sz, C.c_int, vp, C.c_float]
        L.rgc_reg_set_vgicp.argtypes = [vp, C.c_int, C.c_double, C.c_int, C.c_int]
        L.rgc_reg_last_inliers.argtypes = [vp, C.POINTER(C.c_int)]
        L.rgc_reg_set_target_covariance_mode.argtypes = [vp, C.c_int]
        L.rgc_ctx_last_ondemand_ms.argtypes = [vp, C.POINTER(C.c_float)]
        L.rgc_reg_get_voxels.argtypes = [vp, vp, vp, vp, vp, sz, C.POINTER(sz)]
        L.rgc_map_create.argtypes = [vp, vp, sz, sz, C.POINTER(vp)]
        L.rgc_map_destroy.argtypes = [vp]
        for name in ("rgc_map_associate_edges", "rgc_map_associate_planes"):
            getattr(L, name).argtypes = [vp, vp, sz, sz, vp, vp, vp, vp, vp, C.POINTER(sz)]
        L.rgc_voxel_grid.argtypes = [vp, vp, sz, sz, sz, C.c_float, vp, sz, C.POINTER(sz), C.POINTER(C.c_int)]
        L.rgc_deskew.argtypes = [vp, vp, sz, sz, sz, vp, vp, C.c_float, vp]
        for name in ("rgc_reg_set_source_filtered", "rgc_reg_set_target_filtered"):
            getattr(L, name).argtypes = [vp, vp, sz, sz, sz, C.c_float, vp, vp, C.c_float, u64, C.POINTER(sz)]
        _LIB = L
    return _LIB


EXPORTED_SYMBOLS = [
    "rgc_ctx_create", "rgc_ctx_destroy", "rgc_last_error", "rgc_ctx_synchronize", "rgc_ctx_stream", "rgc_ctx_launch_count",
    "rgc_reg_create", "rgc_reg_destroy", "rgc_params_default", "rgc_reg_set_params", "rgc_reg_get_params",
    "rgc_reg_set_source", "rgc_reg_set_target", "rgc_reg_set_source_device", "rgc_reg_set_target_device",
    "rgc_reg_swap_source_and_target", "rgc_reg_clear_source", "rgc_reg_clear_target",
    "rgc_reg_set_source_covs", "rgc_reg_set_target_covs", "rgc_reg_get_source_covs", "rgc_reg_get_target_covs",
    "rgc_reg_align", "rgc_reg_linearize", "rgc_reg_compute_error", "rgc_reg_get_correspondences", "rgc_reg_fitness",
    "rgc_reg_get_final_transformation", "rgc_knn", "rgc_reg_stage_ms", "rgc_knn_self", "rgc_reg_set_owner_slab", "rgc_reg_set_allreduce",
    "rgc_ctx_set_profiling", "rgc_ctx_last_kernel_ms", "rgc_reg_set_vgicp", "rgc_reg_get_voxels", "rgc_reg_last_inliers",
    "rgc_voxel_grid", "rgc_deskew", "rgc_reg_set_source_filtered", "rgc_reg_set_target_filtered",
    "rgc_map_create", "rgc_map_destroy", "rgc_map_associate_edges", "rgc_map_associate_planes",
    "rgc_reg_set_target_covariance_mode", "rgc_ctx_last_ondemand_ms",
    "rgc_batch_align", "rgc_batch_last_stage_ms", "rgc_reg_sync_inputs",
    "rgc_reg_set_target_slab", "rgc_comm_unique_id", "rgc_comm_create", "rgc_comm_destroy", "rgc_comm_info", "rgc_reg_set_comm", "rgc_comm_allreduce_us",
]


class Context:
    """rgc_ctx: one CUDA device + stream + pooled device memory.  No CPU fallback."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        rc = lib().rgc_ctx_create(device, C.byref(self._h))
        if rc != 0:
            raise RgcError(f"rgc_ctx_create(device={device}) failed with {rc}: no usable CUDA device (there is no CPU fallback)")
        self.device = device
        self._regs = weakref.WeakSet()  # registration objects that must be destroyed before the context

    def check(self, rc: int):
        if rc != 0:
            raise RgcError(f"rgc error {rc}: {lib().rgc_last_error(self._h).decode()}")

    @property
    def stream(self) -> int:
        return lib().rgc_ctx_stream(self._h)

    @property
    def launch_count(self) -> int:
        return int(lib().rgc_ctx_launch_count(self._h))

    def synchronize(self):
        self.check(lib().rgc_ctx_synchronize(self._h))

    def set_profiling(self, on: bool):
        self.check(lib().rgc_ctx_set_profiling(self._h, int(on)))

    def last_kernel_ms(self):
        ms = np.zeros(3, np.float32)
        lib().rgc_ctx_last_kernel_ms(self._h, ms.ctypes.data)
        d = dict(zip(("k_correspond", "k_linearize", "k_compute_error"), ms.tolist()))
        od = C.c_float(0.0)
        lib().rgc_ctx_last_ondemand_ms(self._h, C.byref(od))
        d["ondemand_knn_cov"] = od.value
        return d

    def close(self):
        if self._h:
            for r in list(self._regs):
                r._destroy()
            lib().rgc_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_DEFAULT_CTX: dict[int, Context] = {}


def default_context(device: int = 0) -> Context:
    if device not in _DEFAULT_CTX:
        _DEFAULT_CTX[device] = Context(device)
    return _DEFAULT_CTX[device]


def _as_cloud(points):
    """-> (pointer, n, stride_bytes, on_device, keepalive).  Accepts (n, >=3) float32 numpy arrays
    (host) or torch CUDA tensors (device-resident)."""
    if hasattr(points, "is_cuda"):
        t = points
        if t.dtype is not __import__("torch").float32 or t.dim() != 2 or t.shape[1] < 3 or t.stride(1) != 1:
            raise ValueError("device clouds must be float32 tensors of shape (n, >=3) with unit inner stride")
        if t.is_cuda:
            return t.data_ptr(), t.shape[0], t.stride(0) * 4, True, t
        # pinned / pageable host tensor
        return t.data_ptr(), t.shape[0], t.stride(0) * 4, False, t
    a = np.asarray(points)
    if a.dtype != np.float32 or a.ndim != 2 or a.shape[1] < 3 or a.strides[1] != 4:
        a = np.ascontiguousarray(a, dtype=np.float32)
        if a.ndim != 2 or a.shape[1] < 3:
            raise ValueError("point clouds must have shape (n, >=3)")
    return a.ctypes.data, a.shape[0], a.strides[0], False, a


class FastGICP:
    """Mirror of ``fast_gicp::FastGICP<PointSource, PointTarget>`` running on one B200."""

    def __init__(self, ctx: Context | None = None, device: int = 0):
        self.ctx = ctx or default_context(device)
        self._h = C.c_void_p()
        self.ctx.check(lib().rgc_reg_create(self.ctx._h, C.byref(self._h)))
        self.ctx._regs.add(self)
        self._p = _Params()
        lib().rgc_params_default(C.byref(self._p))
        self._src = self._tgt = None
        self._src_id = self._tgt_id = None
        self._n_src = self._n_tgt = 0
        self._final = np.eye(4, dtype=np.float32)
        self._result = None
        self.output = None

    def _destroy(self):
        if self._h and self.ctx._h:
            lib().rgc_reg_destroy(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    # ---- parameter setters (pcl::Registration + fast_gicp names) ----
    def _push(self):
        self.ctx.check(lib().rgc_reg_set_params(self._h, C.byref(self._p)))

    def setMaximumIterations(self, n):
        self._p.max_iterations = int(n); self._push()

    def setMaxCorrespondenceDistance(self, d):
        self._p.max_correspondence_distance = float(d); self._push()

    def setTransformationEpsilon(self, e):
        self._p.transformation_epsilon = float(e); self._push()

    def setRotationEpsilon(self, e):
        self._p.rotation_epsilon = float(e); self._push()

    def setEuclideanFitnessEpsilon(self, e):  # no-op for LsqRegistration
        pass

    def setRANSACIterations(self, n):  # no-op for LsqRegistration
        pass

    def setNumThreads(self, n):  # OpenMP knob; meaningless on the GPU, accepted and ignored
        pass

    def setCorrespondenceRandomness(self, k):
        self._p.k_correspondences = int(k); self._push()

    def setRegularizationMethod(self, m):
        self._p.regularization = int(m); self._push()

    def setInitialLambdaFactor(self, f):
        self._p.lm_init_lambda_factor = float(f); self._push()

    def setLMMaxIterations(self, n):
        self._p.lm_max_iterations = int(n); self._push()

    def setOptimizer(self, o):
        self._p.optimizer = int(o); self._push()

    def setDebugPrint(self, on):
        self._p.lm_debug_print = int(bool(on)); self._push()

    def setGridCell(self, s):
        self._p.grid_cell = float(s); self._push()

    def setTargetCovarianceMode(self, on_demand: bool):
        """True (default): target covariances are computed the first time a target point becomes a
        correspondence; False: all of them at the first align, the reference's schedule
        (fast_gicp_impl.hpp:107-109).  linearize sees bit-identical values either way."""
        self.ctx.check(lib().rgc_reg_set_target_covariance_mode(self._h, int(bool(on_demand))))

    # ---- clouds ----
    def _set(self, which, cloud, force=False):
        ptr, n, stride, on_dev, keep = _as_cloud(cloud)
        # object identity plays the shared_ptr identity (fast_gicp_impl.hpp:73,84); we keep a
        # reference to the object so its id() cannot be recycled while it is the current cloud.
        # Python has no way to say "same buffer, new contents" (the reference's callers allocate a new
        # cloud per frame): `force=True` rebuilds regardless, for callers that refill a buffer in place.
        key = 0 if force else id(cloud)
        if on_dev:
            # the library reads the tensor on ITS stream: order it after whatever torch has queued
            import torch
            torch.cuda.current_stream(keep.device).synchronize()
        fn = getattr(lib(), f"rgc_reg_set_{which}" + ("_device" if on_dev else ""))
        self.ctx.check(fn(self._h, ptr, n, stride, key))
        if which == "source":
            self._src, self._src_id, self._n_src = keep, (key, cloud), n
        else:
            self._tgt, self._tgt_id, self._n_tgt = keep, (key, cloud), n

    def _set_filtered(self, which, xyzi, leaf, q_wxyz, t, scan_period, force=False):
        """[de-skew] -> [pcl::VoxelGrid] -> setInput*, on the device (include/rgc_preprocess.h).
        xyzi: host [n, 4] float32 (x, y, z, intensity).  Returns the size of the filtered cloud."""
        P = np.ascontiguousarray(xyzi, np.float32)
        if P.ndim != 2 or P.shape[1] != 4:
            raise RgcError("filtered input must be [n, 4] float32 (x, y, z, intensity)")
        q = None if q_wxyz is None else np.ascontiguousarray(q_wxyz, np.float64)
        tt = None if t is None else np.ascontiguousarray(t, np.float64)
        m = C.c_size_t(0)
        fn = getattr(lib(), f"rgc_reg_set_{which}_filtered")
        self.ctx.check(fn(self._h, P.ctypes.data, len(P), 16, 12, float(leaf), None if q is None else q.ctypes.data,
                          None if tt is None else tt.ctypes.data, float(scan_period), 0 if force else id(xyzi), C.byref(m)))
        if which == "source":
            self._src, self._src_id, self._n_src = P, (id(xyzi), xyzi), m.value
        else:
            self._tgt, self._tgt_id, self._n_tgt = P, (id(xyzi), xyzi), m.value
        return m.value

    def setInputSourceFiltered(self, xyzi, leaf, q_last_curr=None, t_last_curr=None, scan_period=0.1, force=False):
        return self._set_filtered("source", xyzi, leaf, q_last_curr, t_last_curr, scan_period, force)

    def setInputTargetFiltered(self, xyzi, leaf, q_last_curr=None, t_last_curr=None, scan_period=0.1, force=False):
        return self._set_filtered("target", xyzi, leaf, q_last_curr, t_last_curr, scan_period, force)

    def setInputSource(self, cloud, force=False):
        """`force=True`: rebuild even if `cloud` is the same Python object as last time (a buffer that
        was refilled in place); the default keeps the reference's pointer-identity caching."""
        self._set("source", cloud, force)

    def setInputTarget(self, cloud, force=False):
        self._set("target", cloud, force)

    def waitInputs(self):
        """Complete the deferred part of setInputSource / setInputTarget now (sort, voxel hash, source
        covariances) and raise if it failed; otherwise the next align / linearize does both."""
        self.ctx.check(lib().rgc_reg_sync_inputs(self._h))

    def swapSourceAndTarget(self):
        self.ctx.check(lib().rgc_reg_swap_source_and_target(self._h))
        self._src, self._tgt = self._tgt, self._src
        self._src_id, self._tgt_id = self._tgt_id, self._src_id
        self._n_src, self._n_tgt = self._n_tgt, self._n_src

    def clearSource(self):
        self.ctx.check(lib().rgc_reg_clear_source(self._h)); self._src = self._src_id = None; self._n_src = 0

    def clearTarget(self):
        self.ctx.check(lib().rgc_reg_clear_target(self._h)); self._tgt = self._tgt_id = None; self._n_tgt = 0

    # covariances: (n, 4, 4) float64, symmetric, row/col 3 zero
    def setSourceCovariances(self, covs):
        c = np.ascontiguousarray(covs, np.float64)
        self.ctx.check(lib().rgc_reg_set_source_covs(self._h, c.ctypes.data, c.shape[0]))

    def setTargetCovariances(self, covs):
        c = np.ascontiguousarray(covs, np.float64)
        self.ctx.check(lib().rgc_reg_set_target_covs(self._h, c.ctypes.data, c.shape[0]))

    def getSourceCovariances(self):
        c = np.empty((self._n_src, 4, 4), np.float64)
        self.ctx.check(lib().rgc_reg_get_source_covs(self._h, c.ctypes.data, self._n_src))
        return c

    def getTargetCovariances(self):
        c = np.empty((self._n_tgt, 4, 4), np.float64)
        self.ctx.check(lib().rgc_reg_get_target_covs(self._h, c.ctypes.data, self._n_tgt))
        return c

    # ---- registration ----
    def align(self, guess=None, want_output=False):
        """pcl::Registration::align(output, guess).  Returns the final 4x4 (float32, row-major)."""
        g = None if guess is None else np.ascontiguousarray(np.asarray(guess, np.float32).T)
        T = np.empty((4, 4), np.float32)
        res = _Result()
        out = np.empty((self._n_src, 4), np.float32) if want_output else None
        self.ctx.check(lib().rgc_reg_align(self._h, None if g is None else g.ctypes.data, T.ctypes.data, C.byref(res),
                                           None if out is None else out.ctypes.data))
        self._final = np.ascontiguousarray(T.T)
        self._result = res
        self.output = out
        return self._final

    def getFinalTransformation(self):
        return self._final

    def hasConverged(self):
        return bool(self._result.converged) if self._result is not None else False

    def getFinalHessian(self):
        return np.array(self._result.final_hessian).reshape(6, 6).T.copy()

    @property
    def last_result(self):
        r = self._result
        return dict(converged=bool(r.converged), iterations=r.iterations, n_linearize=r.n_linearize, n_compute_error=r.n_compute_error,
                    n_inliers=r.n_inliers, final_error=r.final_error, device_ms=r.device_ms)

    def last_inliers(self):
        """number of correspondences the last linearize used"""
        n = C.c_int()
        self.ctx.check(lib().rgc_reg_last_inliers(self._h, C.byref(n)))
        return n.value

    def getFitnessScore(self, max_range=np.finfo(np.float64).max):
        s = C.c_double()
        self.ctx.check(lib().rgc_reg_fitness(self._h, float(max_range), C.byref(s)))
        return s.value

    def evaluateCost(self, relative_pose, want_Hb=False):
        """LsqRegistration::evaluateCost: the pose is cast to float first (lsq_registration_impl.hpp:49-50)."""
        T = np.asarray(relative_pose, np.float32).astype(np.float64)
        return self.linearize(T, want_Hb)

    def linearize(self, T, want_Hb=True):
        Tc = np.ascontiguousarray(np.asarray(T, np.float64).T)
        err = C.c_double()
        if want_Hb:
            H, b = np.empty((6, 6)), np.empty(6)
            self.ctx.check(lib().rgc_reg_linearize(self._h, Tc.ctypes.data, C.byref(err), H.ctypes.data, b.ctypes.data))
            return err.value, H.T.copy(), b
        self.ctx.check(lib().rgc_reg_linearize(self._h, Tc.ctypes.data, C.byref(err), None, None))
        return err.value

    def compute_error(self, T):
        Tc = np.ascontiguousarray(np.asarray(T, np.float64).T)
        err = C.c_double()
        self.ctx.check(lib().rgc_reg_compute_error(self._h, Tc.ctypes.data, C.byref(err)))
        return err.value

    def target_cov_state(self):
        """test hook: (covariances (n, 4, 4) as they currently are on the device, computed-flag (n,)),
        without triggering any computation"""
        L = lib()
        L.rgc_debug_get_target_cov_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        c = np.empty((self._n_tgt, 4, 4), np.float64)
        st = np.empty(self._n_tgt, np.int32)
        self.ctx.check(L.rgc_debug_get_target_cov_state(self._h, c.ctypes.data, st.ctypes.data))
        return c, st

    def correspondences(self):
        corr = np.empty(self._n_src, np.int32)
        d2 = np.empty(self._n_src, np.float32)
        self.ctx.check(lib().rgc_reg_get_correspondences(self._h, corr.ctypes.data, d2.ctypes.data))
        return corr, d2

    def stage_ms(self):
        ms = np.zeros(7, np.float32)
        lib().rgc_reg_stage_ms(self._h, ms.ctypes.data)
        return dict(zip(("src_build", "src_knn", "src_cov", "tgt_build", "tgt_knn", "tgt_cov", "lm"), ms.tolist()))


DIRECT27, DIRECT7, DIRECT1 = 0, 1, 2                      # fast_gicp::NeighborSearchMethod (gicp_settings.hpp:8)
ADDITIVE, ADDITIVE_WEIGHTED, MULTIPLICATIVE = 0, 1, 2     # fast_gicp::VoxelAccumulationMode (gicp_settings.hpp:10)


class FastVGICP(FastGICP):
    """Mirror of ``fast_gicp::FastVGICP`` (fast_vgicp.hpp:24-80): the voxelised variant the odometry node
    instantiates (RGC_odometer.cpp:998-1011)."""

    def __init__(self, ctx: Context | None = None, device: int = 0):
        super().__init__(ctx, device)
        self._res, self._search, self._mode = 1.0, DIRECT1, ADDITIVE
        self._push_vox()

    def _push_vox(self):
        self.ctx.check(lib().rgc_reg_set_vgicp(self._h, 1, self._res, self._search, self._mode))

    def setResolution(self, r):
        self._res = float(r); self._push_vox()

    def setNeighborSearchMethod(self, m):
        self._search = int(m); self._push_vox()

    def setVoxelAccumulationMode(self, m):
        self._mode = int(m); self._push_vox()

    def voxels(self):
        """(coords (nv,3) int32, num_points, mean (nv,3), cov upper triangle (nv,6)) sorted by coords."""
        nv = C.c_size_t()
        self.ctx.check(lib().rgc_reg_get_voxels(self._h, None, None, None, None, 0, C.byref(nv)))
        n = nv.value
        coords, num = np.zeros((n, 3), np.int32), np.zeros(n, np.int32)
        mean, cov = np.zeros((n, 3)), np.zeros((n, 6))
        self.ctx.check(lib().rgc_reg_get_voxels(self._h, coords.ctypes.data, num.ctypes.data, mean.ctypes.data, cov.ctypes.data, n, C.byref(nv)))
        o = np.lexsort((coords[:, 2], coords[:, 1], coords[:, 0]))
        return coords[o], num[o], mean[o], cov[o]


def knn(points, queries, k, ctx: Context | None = None, grid_cell: float = 0.0):
    """Exact kNN (pcl::search::KdTree::nearestKSearch) -> (idx int32 (m,k), d2 float32 (m,k))."""
    ctx = ctx or default_context(0)
    pp, n, ps, pd, keep1 = _as_cloud(points)
    qp, m, qs, qd, keep2 = _as_cloud(queries)
    if pd or qd:
        raise ValueError("knn() takes host arrays")
    idx = np.empty((m, k), np.int32)
    d2 = np.empty((m, k), np.float32)
    ctx.check(lib().rgc_knn(ctx._h, pp, n, ps, qp, m, qs, k, idx.ctypes.data, d2.ctypes.data, grid_cell))
    return idx, d2


class FeatureMap:
    """kdtree*FromMap of the mapping node (RGC_mapping.cpp:1073-1074) + the association loops that query it
    (include/rgc_mapping.h).  points: [n, >=3] float32 map points; features: [m, >=3] float32."""

    def __init__(self, points, ctx: Context | None = None):
        self.ctx = ctx or default_context(0)
        P = np.ascontiguousarray(points, np.float32)
        self._h = C.c_void_p()
        self.ctx.check(lib().rgc_map_create(self.ctx._h, P.ctypes.data, P.shape[0], P.strides[0], C.byref(self._h)))
        self.n = P.shape[0]

    def close(self):
        if self._h and self.ctx._h:
            lib().rgc_map_destroy(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _assoc(self, fn, feats, q_wxyz, t, plane):
        F = np.ascontiguousarray(feats, np.float32)
        q, tt = np.ascontiguousarray(q_wxyz, np.float64), np.ascontiguousarray(t, np.float64)
        n = F.shape[0]
        valid = np.zeros(n, np.int32)
        o1 = np.zeros((n, 3), np.float64)
        o2 = np.zeros(n if plane else (n, 3), np.float64)
        nv = C.c_size_t(0)
        self.ctx.check(fn(self._h, F.ctypes.data, n, F.strides[0], q.ctypes.data, tt.ctypes.data, valid.ctypes.data, o1.ctypes.data, o2.ctypes.data, C.byref(nv)))
        return valid.astype(bool), o1, o2

    def associate_edges(self, feats, q_w_curr_wxyz, t_w_curr):
        """-> (valid, point_a, point_b): the arguments of LidarEdgeFactor::Create per feature"""
        return self._assoc(lib().rgc_map_associate_edges, feats, q_w_curr_wxyz, t_w_curr, False)

    def associate_planes(self, feats, q_w_curr_wxyz, t_w_curr):
        """-> (valid, norm, negative_OA_dot_norm): the arguments of LidarPlaneNormFactor::Create per feature"""
        return self._assoc(lib().rgc_map_associate_planes, feats, q_w_curr_wxyz, t_w_curr, True)


def voxel_grid(xyzi, leaf, ctx: Context | None = None):
    """pcl::VoxelGrid<PointXYZI>::filter on the GPU: [n, 4] float32 -> [m, 4] centroids in voxel-index order"""
    ctx = ctx or default_context(0)
    P = np.ascontiguousarray(xyzi, np.float32)
    out = np.empty_like(P)
    m, pt = C.c_size_t(0), C.c_int(0)
    ctx.check(lib().rgc_voxel_grid(ctx._h, P.ctypes.data, len(P), 16, 12, float(leaf), out.ctypes.data, len(P), C.byref(m), C.byref(pt)))
    return out[:m.value].copy()


def deskew(xyzi, q_last_curr_wxyz, t_last_curr, scan_period=0.1, ctx: Context | None = None):
    """RGC_odometer::adjustDistortion for one cloud on the GPU: [n, 4] float32 -> [n, 4]"""
    ctx = ctx or default_context(0)
    P = np.ascontiguousarray(xyzi, np.float32)
    out = np.empty_like(P)
    q = np.ascontiguousarray(q_last_curr_wxyz, np.float64)
    t = np.ascontiguousarray(t_last_curr, np.float64)
    ctx.check(lib().rgc_deskew(ctx._h, P.ctypes.data, len(P), 16, 12, q.ctypes.data, t.ctypes.data, float(scan_period), out.ctypes.data))
    return out


def knn_self(points, k, ctx: Context | None = None, grid_cell: float = 0.0):
    """kNN of every point within its own cloud through the production (tile) kernel -> idx (n, k)."""
    ctx = ctx or default_context(0)
    pp, n, ps, pd, keep = _as_cloud(points)
    idx = np.empty((n, k), np.int32)
    ctx.check(lib().rgc_knn_self(ctx._h, pp, n, ps, k, idx.ctypes.data, grid_cell))
    return idx
