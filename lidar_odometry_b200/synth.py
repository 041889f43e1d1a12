"""Seeded synthetic LiDAR sequences shaped like the reference's datasets.

* ``kitti_sequence``  — HDL-64-like spinning sensor (64 rings, -24.8..+2.0 deg, 1900 azimuth steps,
  ~121 600 returns/scan before drop-outs) driving down a procedural street: ground plane z=-1.73 m,
  building boxes on both sides, parked-car boxes and poles.  Points are stored ring-major then
  azimuth as float32 ``xyzI`` (16 B) exactly like KITTI ``.bin`` files that
  ``util::load_kitti_binary`` reads (/root/reference/src/util/PointCloudUtils.cpp:19-65).
* ``mid360_sequence`` — MID360-like non-repetitive rosette, 20 000 pts/scan, FOV 360 x [-7, +52] deg,
  inside a hall with pillars and crates (<= 40 m).

Everything is a pure function of the integer seed.  No file from /root/reference is read.
"""
from __future__ import annotations

import numpy as np

__all__ = ["Scene", "kitti_scene", "kitti_sequence", "mid360_scene", "mid360_sequence", "pose_matrix"]


class Scene:
    """Axis-aligned boxes (hit from outside), one enclosing room (hit from inside, optional),
    vertical cylinders and a ground plane."""

    def __init__(self, ground_z=None, boxes=None, cylinders=None, room=None):
        self.ground_z = ground_z
        self.boxes = np.asarray(boxes if boxes is not None else np.zeros((0, 6)), dtype=np.float64).reshape(-1, 6)
        self.cylinders = np.asarray(cylinders if cylinders is not None else np.zeros((0, 5)), dtype=np.float64).reshape(-1, 5)
        self.room = None if room is None else np.asarray(room, dtype=np.float64)

    def cast(self, o, d, max_range):
        """o: (3,) origin, d: (N,3) unit directions (world).  Returns ranges (inf = no hit)."""
        n = d.shape[0]
        t = np.full(n, np.inf)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / d
            if self.ground_z is not None:
                tg = (self.ground_z - o[2]) * inv[:, 2]
                tg[~(tg > 1e-6)] = np.inf
                t = np.minimum(t, tg)
            if self.room is not None:
                lo, hi = self.room[:3], self.room[3:]
                t1 = (lo - o) * inv
                t2 = (hi - o) * inv
                tfar = np.min(np.maximum(t1, t2), axis=1)
                tfar[~(tfar > 1e-6)] = np.inf
                t = np.minimum(t, tfar)
            if len(self.boxes):
                c = 0.5 * (self.boxes[:, :3] + self.boxes[:, 3:])
                rad = 0.5 * np.linalg.norm(self.boxes[:, 3:] - self.boxes[:, :3], axis=1)
                near = np.linalg.norm(c - o, axis=1) - rad < max_range
                for b in self.boxes[near]:
                    t1 = (b[:3] - o) * inv
                    t2 = (b[3:] - o) * inv
                    tn = np.max(np.minimum(t1, t2), axis=1)
                    tf = np.min(np.maximum(t1, t2), axis=1)
                    hit = (tn <= tf) & (tn > 1e-6)
                    t = np.where(hit & (tn < t), tn, t)
            for cx, cy, r, z0, z1 in self.cylinders:
                if np.hypot(cx - o[0], cy - o[1]) - r > max_range:
                    continue
                ox, oy = o[0] - cx, o[1] - cy
                a = d[:, 0] ** 2 + d[:, 1] ** 2
                bq = 2.0 * (ox * d[:, 0] + oy * d[:, 1])
                cq = ox * ox + oy * oy - r * r
                disc = bq * bq - 4 * a * cq
                ok = (disc > 0) & (a > 1e-12)
                tc = (-bq - np.sqrt(np.where(ok, disc, 0.0))) / (2 * np.where(ok, a, 1.0))
                z = o[2] + tc * d[:, 2]
                hit = ok & (tc > 1e-6) & (z >= z0) & (z <= z1)
                t = np.where(hit & (tc < t), tc, t)
        return t


def _cast_torch(scene, o, d, max_range, device):
    """Same ray casting as Scene.cast, evaluated with torch (float64) on `device` - used by bench.py to build
    full-size sequences quickly.  Plumbing for synthetic data only; not part of the registration path."""
    import torch
    dt = torch.float64
    d_t = torch.as_tensor(d, dtype=dt, device=device)
    o_t = torch.as_tensor(np.asarray(o, np.float64), dtype=dt, device=device)
    inf = torch.tensor(float("inf"), dtype=dt, device=device)
    n = d_t.shape[0]
    t = torch.full((n,), float("inf"), dtype=dt, device=device)
    inv = 1.0 / d_t
    if scene.ground_z is not None:
        tg = (scene.ground_z - o_t[2]) * inv[:, 2]
        tg = torch.where(tg > 1e-6, tg, inf)
        t = torch.minimum(t, tg)
    if scene.room is not None:
        lo = torch.as_tensor(scene.room[:3], dtype=dt, device=device); hi = torch.as_tensor(scene.room[3:], dtype=dt, device=device)
        t1 = (lo - o_t) * inv; t2 = (hi - o_t) * inv
        tfar = torch.maximum(t1, t2).min(dim=1).values
        tfar = torch.where(tfar > 1e-6, tfar, inf)
        t = torch.minimum(t, tfar)
    if len(scene.boxes):
        c = 0.5 * (scene.boxes[:, :3] + scene.boxes[:, 3:])
        rad = 0.5 * np.linalg.norm(scene.boxes[:, 3:] - scene.boxes[:, :3], axis=1)
        near = np.linalg.norm(c - np.asarray(o), axis=1) - rad < max_range
        for b in scene.boxes[near]:
            bt = torch.as_tensor(b, dtype=dt, device=device)
            t1 = (bt[:3] - o_t) * inv; t2 = (bt[3:] - o_t) * inv
            tn = torch.minimum(t1, t2).max(dim=1).values
            tf = torch.maximum(t1, t2).min(dim=1).values
            hit = (tn <= tf) & (tn > 1e-6) & (tn < t)
            t = torch.where(hit, tn, t)
    for cx, cy, r, z0, z1 in scene.cylinders:
        if np.hypot(cx - o[0], cy - o[1]) - r > max_range:
            continue
        ox, oy = float(o[0] - cx), float(o[1] - cy)
        a = d_t[:, 0] ** 2 + d_t[:, 1] ** 2
        bq = 2.0 * (ox * d_t[:, 0] + oy * d_t[:, 1])
        cq = ox * ox + oy * oy - r * r
        disc = bq * bq - 4 * a * cq
        ok = (disc > 0) & (a > 1e-12)
        tc = (-bq - torch.sqrt(torch.where(ok, disc, torch.zeros_like(disc)))) / (2 * torch.where(ok, a, torch.ones_like(a)))
        z = o_t[2] + tc * d_t[:, 2]
        hit = ok & (tc > 1e-6) & (z >= z0) & (z <= z1) & (tc < t)
        t = torch.where(hit, tc, t)
    return t.cpu().numpy()


def pose_matrix(x, y, z, yaw, pitch=0.0, roll=0.0):
    cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = (x, y, z)
    return T


def kitti_scene(seed=42, length=420.0, x0=-140.0):
    rng = np.random.default_rng(seed)
    boxes, cyl = [], []
    for side in (-1.0, 1.0):
        x = x0
        while x < x0 + length:
            w = rng.uniform(8.0, 25.0)
            depth = rng.uniform(8.0, 15.0)
            h = rng.uniform(4.0, 18.0)
            setback = rng.uniform(9.0, 13.0)
            y_in = side * setback
            y_out = side * (setback + depth)
            boxes.append([x, min(y_in, y_out), -1.73, x + w, max(y_in, y_out), -1.73 + h])
            x += w + rng.uniform(0.0, 6.0)
        # parked cars
        x = x0 + rng.uniform(0, 10)
        while x < x0 + length:
            if rng.uniform() < 0.6:
                yc = side * rng.uniform(5.0, 6.0)
                boxes.append([x, yc - 0.9, -1.73, x + rng.uniform(3.8, 4.8), yc + 0.9, -1.73 + rng.uniform(1.3, 1.7)])
            x += rng.uniform(6.0, 14.0)
        # poles / trunks
        x = x0 + rng.uniform(0, 15)
        while x < x0 + length:
            cyl.append([x, side * rng.uniform(7.0, 8.0), rng.uniform(0.12, 0.3), -1.73, -1.73 + rng.uniform(4.0, 9.0)])
            x += rng.uniform(12.0, 30.0)
    return Scene(ground_z=-1.73, boxes=boxes, cylinders=cyl)


def _hdl64_dirs(n_rings=64, n_az=1900):
    elev = np.deg2rad(np.linspace(2.0, -24.8, n_rings))
    az = np.linspace(0.0, 2 * np.pi, n_az, endpoint=False)
    ce, se = np.cos(elev)[:, None], np.sin(elev)[:, None]
    d = np.stack([ce * np.cos(az)[None, :], ce * np.sin(az)[None, :], np.broadcast_to(se, (n_rings, n_az))], axis=-1)
    return d.reshape(-1, 3)


def kitti_trajectory(n_scans, step=1.2, seed=42):
    """~1.2 m/scan forward with slow yaw (<= 2 deg/scan): a keyframe on nearly every scan."""
    rng = np.random.default_rng(seed + 1000)
    poses = []
    x = y = 0.0
    yaw = 0.0
    for k in range(n_scans):
        poses.append(pose_matrix(x, y, 0.0, yaw, pitch=0.002 * np.sin(0.3 * k), roll=0.002 * np.cos(0.23 * k)))
        yaw_rate = np.deg2rad(0.8) * np.sin(2 * np.pi * k / 60.0) + np.deg2rad(0.1) * rng.standard_normal()
        yaw += float(np.clip(yaw_rate, -np.deg2rad(2.0), np.deg2rad(2.0)))
        s = step * (1.0 + 0.03 * rng.standard_normal())
        x += s * np.cos(yaw)
        y += s * np.sin(yaw)
    return poses


def _scan(scene, T, dirs_local, max_range, sigma, rng, dropout=0.0, device=None):
    o = T[:3, 3]
    d = dirs_local @ T[:3, :3].T
    t = scene.cast(o, d, max_range) if device is None else _cast_torch(scene, o, d, max_range, device)
    keep = np.isfinite(t) & (t < max_range) & (t > 0.5)
    if dropout > 0:
        keep &= rng.uniform(size=t.shape) >= dropout
    r = t[keep] + sigma * rng.standard_normal(int(keep.sum()))
    pts = dirs_local[keep] * r[:, None]
    out = np.empty((pts.shape[0], 4), dtype=np.float32)
    out[:, :3] = pts.astype(np.float32)
    out[:, 3] = rng.uniform(0.0, 1.0, pts.shape[0]).astype(np.float32)
    return out


def kitti_sequence(n_scans=100, seed=42, n_rings=64, n_az=1900, max_range=100.0, sigma=0.02, device=None):
    """Returns (list of (N_i,4) float32 xyzI scans in the sensor frame, list of 4x4 ground-truth poses)."""
    scene = kitti_scene(seed)
    dirs = _hdl64_dirs(n_rings, n_az)
    poses = kitti_trajectory(n_scans, seed=seed)
    rng = np.random.default_rng(seed + 7)
    scans = [_scan(scene, T, dirs, max_range, sigma, rng, dropout=0.01, device=device) for T in poses]
    return scans, poses


def mid360_scene(seed=42):
    rng = np.random.default_rng(seed)
    room = [-20.0, -12.0, -1.2, 20.0, 12.0, 6.0]
    boxes, cyl = [], []
    for _ in range(14):
        cx, cy = rng.uniform(-18, 18), rng.uniform(-10, 10)
        if abs(cx) < 3 and abs(cy) < 3:
            continue
        sx, sy, sz = rng.uniform(0.6, 2.5), rng.uniform(0.6, 2.5), rng.uniform(0.8, 3.0)
        boxes.append([cx - sx / 2, cy - sy / 2, -1.2, cx + sx / 2, cy + sy / 2, -1.2 + sz])
    for _ in range(8):
        cx, cy = rng.uniform(-18, 18), rng.uniform(-10, 10)
        if abs(cx) < 3 and abs(cy) < 3:
            continue
        cyl.append([cx, cy, rng.uniform(0.2, 0.5), -1.2, 6.0])
    return Scene(ground_z=None, boxes=boxes, cylinders=cyl, room=room)


def _mid360_dirs(n, k, seed):
    """Non-repetitive pattern: a rank-1 lattice whose phase advances with the scan index."""
    g1, g2 = 0.7548776662466927, 0.5698402909980532  # plastic-number lattice (low discrepancy)
    i = np.arange(n, dtype=np.float64) + 1.0
    u = (i * g1 + 0.137 * k + 0.01 * seed) % 1.0
    v = (i * g2 + 0.311 * k) % 1.0
    az = 2 * np.pi * u
    lo, hi = np.sin(np.deg2rad(-7.0)), np.sin(np.deg2rad(52.0))
    se = lo + (hi - lo) * v
    ce = np.sqrt(1 - se * se)
    return np.stack([ce * np.cos(az), ce * np.sin(az), se], axis=-1)


def mid360_sequence(n_scans=100, seed=42, n_pts=20000, max_range=40.0, sigma=0.01, device=None):
    scene = mid360_scene(seed)
    rng = np.random.default_rng(seed + 7)
    scans, poses = [], []
    for k in range(n_scans):
        ang = 2 * np.pi * k / 240.0
        T = pose_matrix(2.5 * np.cos(ang) - 2.5 + 0.1 * 0, 2.5 * np.sin(ang), 0.0, yaw=ang * 0.5)
        # ~0.1 m / scan along a slow arc
        T[:3, 3] = (0.1 * k * np.cos(0.2 * ang), 0.1 * k * np.sin(0.2 * ang) * 0.3, 0.0)
        dirs = _mid360_dirs(n_pts, k, seed)
        scans.append(_scan(scene, T, dirs, max_range, sigma, rng, device=device))
        poses.append(T)
    return scans, poses
