"""Debug aid (GPU box): run the oracle pipeline and the CUDA odometry side by side and print where they part."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import orc
from lidar_odometry_b200 import api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
scans, poses = synth.kitti_sequence(n_scans=n, seed=7, n_rings=64, n_az=600)
pipe = orc.Pipeline(); odo = api.Odometry()
icp = api.IterativeClosestPointOptimizer(api.ICPConfig(), api.AdaptiveMEstimator())
for k, s in enumerate(scans):
    # teacher-forced check BEFORE stepping: run both ICPs from the oracle's state
    a = pipe.process(s); b = odo.process(s)
    dt = np.linalg.norm(a["pose"][:3, 3].astype(np.float64) - b["pose"][:3, 3])
    Ra, Rb = a["pose"][:3, :3].astype(np.float64), b["pose"][:3, :3].astype(np.float64)
    A = Ra.T @ Rb; ang = np.linalg.norm(0.5 * np.array([A[2, 1] - A[1, 2], A[0, 2] - A[2, 0], A[1, 0] - A[0, 1]]))
    print(f"scan {k}: feat {a['n_features']}/{b['n_features']} kf {a['keyframe']}/{b['keyframe']} ok {a['icp_ok']}/{b['icp_ok']} "
          f"corr {a['n_corr']}/{b['n_corr']} iters {a['n_iters']}/{b['n_iters']} dpos {dt:.3e} drot {ang:.3e} L0 {pipe.map().counts()[0]}/{b['l0']} dev_ms {b['device_ms']:.3f}")
