import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rgc_slam_b200 import synth
from rgc_slam_b200.features import extract_features
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
scene = synth.Scene.make(synth.BASE_SEED + 3000)
traj = synth.trajectory(12, seed=3)
scans = [synth.lidar_scan(scene, traj[f], seed=synth.BASE_SEED + 3000 + f) for f in range(8)]
batch = [scans[i % 8] for i in range(B)]
extract_features(batch[:4], want_arrays=False)
r, ms = extract_features(batch, want_arrays=False)
print("device_ms", ms)
