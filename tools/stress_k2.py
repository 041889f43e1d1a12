"""GPU box: K2 (surfel correspondence) on the 10^7-voxel map with 2^20 queries - the configuration the roofline is quoted on.
Used under ncu (--set full -k regex:k_icp_corr) to read DRAM traffic, pipe utilisation and stall reasons."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lidar_odometry_b200 import api
rng = np.random.default_rng(1234)
side, layers = 440, 52
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
nq = 1 << 20
q = np.stack([rng.uniform(-109, 109, nq), rng.uniform(-109, 109, nq), -39.0 + 1.5 * rng.integers(0, layers, nq) + 0.7 + rng.normal(0, 0.02, nq)], axis=1).astype(np.float32)
icp = api.IterativeClosestPointOptimizer(api.ICPConfig(max_iterations=1), api.AdaptiveMEstimator())
for _ in range(3):
    ok, T = icp.optimize(vmap, q, np.eye(4, dtype=np.float32))
print("ok", ok, vmap.GetVoxelCount(), icp.get_last_stats().num_correspondences)
