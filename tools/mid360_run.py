"""GPU box: a short MID360-shaped KDTree-mode sequence (used under ncu for the K3 launch list)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lidar_odometry_b200 import api, synth
scans, _ = synth.mid360_sequence(n_scans=16, seed=42, device="cuda")
os.environ["B2LO_NO_GRAPH"] = "1"
odo = api.Odometry(mid360=True)
for s in scans:
    r = odo.process(np.ascontiguousarray(s))
print(r["n_features"], r["n_corr"], r["n_iters"], r["l0"])
