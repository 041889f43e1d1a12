"""Debug aid (GPU box): wall time of the stand-alone b2lo_map_update / b2lo_map_export_l0 calls (the class-by-class drop-in path)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lidar_odometry_b200 import api, synth
from oracle import orc
scans, poses = synth.kitti_sequence(n_scans=60, seed=42, device="cuda")
feats = [orc.voxel_filter(s[:, :3], 8, 0.5)[0] for s in scans]
for hint in (0, 1 << 17):
    ctx = api.Context(0)
    m = api.VoxelMap(0.5, ctx, capacity_hint=hint)
    ts = []
    for k, f in enumerate(feats):
        T = poses[k].astype(np.float32)
        w = (f @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
        ctx.sync(); t0 = time.perf_counter()
        m.UpdateVoxelMap(w, T[:3, 3].astype(np.float64), 120.0)
        ctx.sync(); t1 = time.perf_counter()
        c = m.GetPointCloud() if hasattr(m, "GetPointCloud") else m.export_l0()[0]
        t2 = time.perf_counter()
        ts.append((1e3 * (t1 - t0), 1e3 * (t2 - t1), len(c)))
    a = np.array(ts)
    print(f"hint {hint}: update ms median {np.median(a[5:,0]):.3f} mean {a[5:,0].mean():.3f} max {a[5:,0].max():.3f}; export ms median {np.median(a[5:,1]):.3f}; L0 {int(a[-1,2])}")
    print("   first 12 update ms:", np.round(a[:12, 0], 3))
