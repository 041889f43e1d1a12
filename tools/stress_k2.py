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
# coherent: a raster sweep over the slabs like a real scan (one query per 0.5 m L0 cell: 9 queries per 1.5 m L1 cell and layer)
ax = np.arange(-109.0, 109.0, 0.5)
xx, yy = np.meshgrid(ax, ax, indexing="ij")
xy = np.stack([xx.ravel(), yy.ravel()], axis=1)
need = (nq + len(xy) - 1) // len(xy)
qc = np.concatenate([np.concatenate([xy, np.full((len(xy), 1), -39.0 + 1.5 * l + 0.7)], axis=1) for l in range(need)])[:nq]
qc = (qc + rng.normal(0, 0.01, qc.shape)).astype(np.float32)
icp = api.IterativeClosestPointOptimizer(api.ICPConfig(max_iterations=1), api.AdaptiveMEstimator())
# launch order of k_icp_corr (what `ncu -k regex:k_icp_corr -s 2 -c 4` picks): random x3, then coherent x3
for name, cloud in (("random", q), ("coherent", qc)):
    for _ in range(3):
        ok, T = icp.optimize(vmap, cloud, np.eye(4, dtype=np.float32))
    print(name, "ok", ok, vmap.GetVoxelCount(), icp.get_last_stats().num_correspondences)
