"""GPU box: ApplyTransformAndRehash on the ~10^7-voxel map (bulk rebuild).  Run under
ncu --metrics gpu__time_duration.sum -k regex:^k_ to list where the rebuild spends its time."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lidar_odometry_b200 import api, synth
rng = np.random.default_rng(1234)
side, layers = 440, int(os.environ.get("LAYERS", "52"))
per_layer = side * side
ctx = api.Context(0)
vmap = api.VoxelMap(0.5, ctx, capacity_hint=int(per_layer * layers * 1.15))
gx, gy = np.meshgrid(np.arange(side, dtype=np.float32), np.arange(side, dtype=np.float32), indexing="ij")
base = np.stack([gx.ravel(), gy.ravel()], axis=1) * np.float32(0.5) - np.float32(110.0)
for l in range(layers):
    pts = np.empty((per_layer, 3), np.float32)
    pts[:, :2] = base + rng.uniform(0.05, 0.45, (per_layer, 2)).astype(np.float32)
    pts[:, 2] = np.float32(-39.0 + 1.5 * l + 0.7) + rng.normal(0.0, 0.01, per_layer).astype(np.float32)
    vmap.UpdateVoxelMap(pts, [0.0, 0.0, 0.0], 400.0)
ctx.sync()
T = synth.pose_matrix(0.31, -0.22, 0.05, 0.02, 0.003, -0.004).astype(np.float32)
print("MARK rehash begins", vmap.GetVoxelCount(), flush=True)
for k in range(2):
    t0 = time.perf_counter()
    vmap.ApplyTransformAndRehash(T if k == 0 else np.linalg.inv(T.astype(np.float64)).astype(np.float32))
    ctx.sync()
    print("rehash", k, vmap.GetVoxelCount(), vmap.GetL1VoxelCount(), vmap.GetSurfelCount(), f"{1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)
