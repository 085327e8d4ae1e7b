"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):
one quarter-resolution align + fitness + correspondences + a 3-scan feature batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rgc_slam_b200 as rgc
from rgc_slam_b200 import synth
from rgc_slam_b200.features import extract_features
scene = synth.Scene.make(synth.BASE_SEED)
traj = synth.trajectory(8, seed=1)
tgt = synth.to_xyz1(synth.lidar_scan(scene, traj[3], n_azimuth=300, seed=1))
src = synth.to_xyz1(synth.lidar_scan(scene, traj[4], n_azimuth=300, seed=2))
g = rgc.FastGICP()
g.setMaxCorrespondenceDistance(2.0)
g.setInputTarget(tgt); g.setInputSource(src)
T = g.align(want_output=True)
print("align", g.last_result, g.getFitnessScore(), (g.correspondences()[0] >= 0).sum())
print("knn", rgc.knn(tgt, src[:500], 5)[0].sum(), rgc.knn_self(tgt, 20).sum())
# the warp-per-query kernel that finishes deferred tiles: force every tile through it once
import ctypes as C
from rgc_slam_b200 import api
L = api.lib()
L.rgc_debug_set_knn_defer.argtypes = [C.c_void_p, C.c_int]
L.rgc_debug_set_knn_defer(api.default_context()._h, 1)
print("knn (deferred path)", rgc.knn_self(tgt, 20).sum())
L.rgc_debug_set_knn_defer(api.default_context()._h, 600)
v = rgc.FastVGICP()
v.setResolution(1.0)
v.setMaxCorrespondenceDistance(2.0)
v.setInputTarget(tgt); v.setInputSource(src)
v.align()
print("vgicp", v.last_result)
scans = [synth.lidar_scan(scene, traj[f], n_azimuth=400, seed=50 + f) for f in range(3)]
r, ms = extract_features(scans)
print("features", [x["cloud_size"] for x in r], [len(x["corner_sharp"]) for x in r])
# batched registration (multi-cloud grids, per-pair reductions, on-demand covariances over a list)
from rgc_slam_b200 import batch
pp = [dict(src=src[: 2000 + 300 * i], tgt=tgt, guess=np.eye(4, dtype=np.float32)) for i in range(3)]
prm = batch.default_params()
prm.max_iterations, prm.max_correspondence_distance = 16, 2.0
res = batch.align_batch(pp, params=prm, want_fitness=True)
print("batch", [(r["iterations"], r["converged"], round(float(r["fitness"]), 4)) for r in res])
