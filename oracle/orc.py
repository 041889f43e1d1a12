"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/liborc.so (the CPU restatement of the reference's hot path, see
oracle/include/orc_*.hpp) and of the pieces of the REAL reference that compile here
(oracle/_ref/*.so).  Imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liborc.so")
REF_PKO = os.path.join(HERE, "_ref", "libref_pko.so")
REF_CONT = os.path.join(HERE, "_ref", "libref_cont.so")


def build(force=False):
    """Compile liborc.so (and oracle/_ref/* when /root/reference is present)."""
    if force or not os.path.exists(LIB) or os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-s", "-C", HERE], check=True)


class IcpCfg(C.Structure):
    _fields_ = [
        ("max_iterations", C.c_int), ("translation_tolerance", C.c_double), ("rotation_tolerance", C.c_double),
        ("max_correspondence_distance", C.c_double), ("min_correspondence_points", C.c_int), ("use_robust_loss", C.c_int),
        ("robust_loss_delta", C.c_double), ("use_surfel_correspondence", C.c_int), ("use_adaptive_m_estimator", C.c_int),
        ("loss_type", C.c_int), ("min_scale_factor", C.c_double), ("max_scale_factor", C.c_double),
        ("num_alpha_segments", C.c_int), ("truncated_threshold", C.c_double), ("gmm_components", C.c_int),
        ("gmm_sample_size", C.c_int), ("pko_kernel_type", C.c_int),
    ]


class IterTrace(C.Structure):
    _fields_ = [
        ("n_corr", C.c_int), ("scale", C.c_double), ("delta", C.c_double),
        ("H", C.c_float * 36), ("g", C.c_float * 6), ("cost", C.c_float),
        ("H64", C.c_double * 36), ("g64", C.c_double * 6), ("cost64", C.c_double),
        ("dx", C.c_float * 6), ("T_in", C.c_float * 16), ("T_out", C.c_float * 16),
        ("em_iters", C.c_int), ("kmeans_iters", C.c_int),
    ]


class PipeCfg(C.Structure):
    _fields_ = [
        ("voxel_size", C.c_float), ("point_stride", C.c_int), ("map_voxel_size", C.c_float), ("max_range", C.c_double),
        ("surfel_planarity_threshold", C.c_float), ("keyframe_distance_threshold", C.c_double),
        ("keyframe_rotation_threshold", C.c_double), ("icp", IcpCfg),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        L.orc_filter_morton_key.restype = C.c_uint64
        L.orc_filter_morton_key.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float]
        L.orc_voxel_key_hash.restype = C.c_uint64
        L.orc_voxel_key_hash.argtypes = [C.c_int, C.c_int, C.c_int]
        L.orc_map_create.restype = C.c_void_p
        L.orc_map_create.argtypes = [C.c_float, C.c_int, C.c_float, C.c_int]
        L.orc_pipe_create.restype = C.c_void_p
        L.orc_pipe_map.restype = C.c_void_p
        L.orc_pipe_map.argtypes = [C.c_void_p]
        L.orc_pipe_features.restype = C.c_size_t
        L.orc_icp_correspondences.restype = C.c_size_t
        L.orc_kdtree_correspondences.restype = C.c_size_t
        L.orc_pko_scale.restype = C.c_double
        _lib = L
    return _lib


def _p(a, t=None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def default_icp_cfg():
    c = IcpCfg()
    lib().orc_default_icp_cfg(C.byref(c))
    return c


def default_pipe_cfg(mid360=False):
    c = PipeCfg()
    lib().orc_default_pipe_cfg(C.byref(c), int(mid360))
    return c


def filter_morton_key(x, y, z, voxel):
    return int(lib().orc_filter_morton_key(x, y, z, voxel))


def voxel_key_hash(x, y, z):
    return int(lib().orc_voxel_key_hash(int(x), int(y), int(z)))


def point_to_key(p, voxel, factor=3, level=0):
    p = f32(p)
    k = np.zeros(3, np.int32)
    lib().orc_point_to_key(_p(p), C.c_float(voxel), int(factor), int(level), _p(k))
    return k


def parent_key(key, factor=3):
    key = np.ascontiguousarray(key, np.int32)
    out = np.zeros(3, np.int32)
    lib().orc_parent_key(_p(key), int(factor), _p(out))
    return out


def voxel_filter(xyz, stride, voxel):
    """xyz: (N,3) float32.  Returns (centroids (M,3) f32, morton keys (M,) u64) in first-seen order."""
    xyz = f32(xyz).reshape(-1, 3)
    n = xyz.shape[0]
    out = np.zeros((max(n, 1), 3), np.float32)
    keys = np.zeros(max(n, 1), np.uint64)
    m = C.c_size_t(0)
    lib().orc_filter(_p(xyz), C.c_size_t(n), int(stride), C.c_float(voxel), _p(out), _p(keys), C.byref(m))
    return out[: m.value].copy(), keys[: m.value].copy()


def _load_image(fn, image):
    a = np.frombuffer(bytes(image), dtype=np.uint8) if not isinstance(image, np.ndarray) else np.ascontiguousarray(image.view(np.uint8)).reshape(-1)
    cap = max(a.size // 4, 1)
    out = np.zeros((cap, 3), np.float32)
    n = C.c_size_t(0)
    rc = fn(_p(a), C.c_size_t(a.size), _p(out), C.c_size_t(cap), C.byref(n))
    assert rc == 0
    return out[: n.value].copy()


def voxel_grid_filter(xyz, leaf):
    """util::VoxelGrid::filter: (N,3) f32 -> centroids (M,3) f32 in std::map key order."""
    xyz = f32(xyz).reshape(-1, 3)
    n = xyz.shape[0]
    out = np.zeros((max(n, 1), 3), np.float32)
    m = C.c_size_t(0)
    rc = lib().orc_voxel_grid_filter(_p(xyz), C.c_size_t(n), C.c_float(leaf), _p(out), C.c_size_t(max(n, 1)), C.byref(m))
    assert rc == 0
    return out[: m.value].copy()


def ply_load(image):
    """PLYPlayer::load_ply_point_cloud on a file image -> (N,3) f32 (empty when the reference rejects the file)."""
    return _load_image(lib().orc_ply_load, image)


def kitti_load(image):
    """util::load_kitti_binary on a file image -> (N,3) f32."""
    return _load_image(lib().orc_kitti_load, image)


class VoxelMap:
    def __init__(self, voxel=0.5, factor=3, planarity=0.1, compute_surfels=True, handle=None):
        self._own = handle is None
        self.h = C.c_void_p(handle) if handle is not None else C.c_void_p(lib().orc_map_create(voxel, factor, planarity, int(compute_surfels)))

    def __del__(self):
        if getattr(self, "_own", False) and self.h:
            lib().orc_map_destroy(self.h)
            self.h = None

    def clear(self):
        lib().orc_map_clear(self.h)

    def update(self, xyz, sensor, max_distance):
        xyz = f32(xyz).reshape(-1, 3)
        s = np.ascontiguousarray(sensor, np.float64)
        lib().orc_map_update(self.h, _p(xyz), C.c_size_t(xyz.shape[0]), _p(s), C.c_double(max_distance))

    def counts(self):
        a, b, c = C.c_size_t(), C.c_size_t(), C.c_size_t()
        lib().orc_map_counts(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def export_l0(self):
        n = self.counts()[0]
        keys = np.zeros((n, 3), np.int32); cent = np.zeros((n, 3), np.float32); cnt = np.zeros(n, np.int32)
        lib().orc_map_export_l0(self.h, _p(keys), _p(cent), _p(cnt))
        return keys, cent, cnt

    def export_l1(self):
        n = self.counts()[1]
        d = dict(keys=np.zeros((n, 3), np.int32), nchild=np.zeros(n, np.int32), children=np.zeros((n, 27, 3), np.int32),
                 has_surfel=np.zeros(n, np.int32), normal=np.zeros((n, 3), np.float32), centroid=np.zeros((n, 3), np.float32),
                 planarity=np.zeros(n, np.float32), last_child_count=np.zeros(n, np.int32))
        lib().orc_map_export_l1(self.h, _p(d["keys"]), _p(d["nchild"]), _p(d["children"]), _p(d["has_surfel"]), _p(d["normal"]),
                                _p(d["centroid"]), _p(d["planarity"]), _p(d["last_child_count"]))
        return d

    def lookup(self, p):
        p = f32(p); n = np.zeros(3, np.float32); c = np.zeros(3, np.float32)
        ok = lib().orc_map_lookup(self.h, _p(p), _p(n), _p(c))
        return bool(ok), n, c

    def transform_rehash(self, T):
        T = f32(T).reshape(16)
        lib().orc_map_transform_rehash(self.h, _p(T))


def icp_correspondences(vmap, local_xyz, T, max_dist=1.0):
    local = f32(local_xyz).reshape(-1, 3)
    m = local.shape[0]
    T = f32(T).reshape(16)
    out = dict(state=np.zeros(m, np.int32), l1key=np.zeros((m, 3), np.int32), morton=np.zeros(m, np.uint64),
               normal=np.zeros((m, 3), np.float32), centroid=np.zeros((m, 3), np.float32), residual=np.zeros(m, np.float64),
               world=np.zeros((m, 3), np.float32))
    n = lib().orc_icp_correspondences(vmap.h, _p(local), C.c_size_t(m), _p(T), C.c_double(max_dist), _p(out["state"]), _p(out["l1key"]),
                                      _p(out["morton"]), _p(out["normal"]), _p(out["centroid"]), _p(out["residual"]), _p(out["world"]))
    out["n_accepted"] = int(n)
    return out


def _trace_to_dicts(tr, n):
    res = []
    for i in range(n):
        t = tr[i]
        res.append(dict(n_corr=t.n_corr, scale=t.scale, delta=t.delta, H=np.array(t.H, np.float32).reshape(6, 6), g=np.array(t.g, np.float32),
                        cost=t.cost, H64=np.array(t.H64).reshape(6, 6), g64=np.array(t.g64), cost64=t.cost64, dx=np.array(t.dx, np.float32),
                        T_in=np.array(t.T_in, np.float32).reshape(4, 4), T_out=np.array(t.T_out, np.float32).reshape(4, 4),
                        em_iters=t.em_iters, kmeans_iters=t.kmeans_iters))
    return res


def icp_optimize(vmap, local_xyz, T_init, cfg=None, trace_cap=64):
    cfg = cfg or default_icp_cfg()
    local = f32(local_xyz).reshape(-1, 3)
    T0 = f32(T_init).reshape(16)
    Tout = np.zeros(16, np.float32)
    tr = (IterTrace * trace_cap)()
    nt = C.c_int(0)
    ok = lib().orc_icp_optimize(vmap.h, _p(local), C.c_size_t(local.shape[0]), _p(T0), C.byref(cfg), _p(Tout), tr, trace_cap, C.byref(nt))
    return bool(ok), Tout.reshape(4, 4), _trace_to_dicts(tr, nt.value)


def icp_optimize_kdtree(map_xyz, local_xyz, T_init, cfg=None, trace_cap=64):
    cfg = cfg or default_icp_cfg()
    mp = f32(map_xyz).reshape(-1, 3)
    local = f32(local_xyz).reshape(-1, 3)
    T0 = f32(T_init).reshape(16)
    Tout = np.zeros(16, np.float32)
    tr = (IterTrace * trace_cap)()
    nt = C.c_int(0)
    ok = lib().orc_icp_optimize_kdtree(_p(mp), C.c_size_t(mp.shape[0]), _p(local), C.c_size_t(local.shape[0]), _p(T0), C.byref(cfg),
                                       _p(Tout), tr, trace_cap, C.byref(nt))
    return bool(ok), Tout.reshape(4, 4), _trace_to_dicts(tr, nt.value)


def icp_optimize_loop(curr_xyz, T_curr, matched_xyz, T_matched, cfg=None, trace_cap=128):
    """optimize_loop (ICP.cpp:40-251) -> (success, T_relative, inlier_ratio, iterations, trace)."""
    cfg = cfg or default_icp_cfg()
    cur = f32(curr_xyz).reshape(-1, 3)
    mat = f32(matched_xyz).reshape(-1, 3)
    Tc, Tm = f32(T_curr).reshape(16), f32(T_matched).reshape(16)
    Trel = np.zeros(16, np.float32)
    ratio, iters, nt = C.c_float(0), C.c_int(0), C.c_int(0)
    tr = (IterTrace * trace_cap)()
    L = lib()
    L.orc_icp_optimize_loop.restype = C.c_int
    ok = L.orc_icp_optimize_loop(_p(cur), C.c_size_t(cur.shape[0]), _p(Tc), _p(mat), C.c_size_t(mat.shape[0]), _p(Tm), C.byref(cfg), _p(Trel),
                                 C.byref(ratio), C.byref(iters), tr, trace_cap, C.byref(nt))
    return bool(ok), Trel.reshape(4, 4), float(ratio.value), int(iters.value), _trace_to_dicts(tr, nt.value)


def kdtree_correspondences(map_xyz, local_xyz, T, max_dist=1.0):
    mp = f32(map_xyz).reshape(-1, 3)
    local = f32(local_xyz).reshape(-1, 3)
    m = local.shape[0]
    T = f32(T).reshape(16)
    out = dict(knn=np.zeros((m, 5), np.int32), state=np.zeros(m, np.int32), normal=np.zeros((m, 3), np.float32),
               centroid=np.zeros((m, 3), np.float32), residual=np.zeros(m, np.float64))
    n = lib().orc_kdtree_correspondences(_p(mp), C.c_size_t(mp.shape[0]), _p(local), C.c_size_t(m), _p(T), C.c_double(max_dist),
                                         _p(out["knn"]), _p(out["state"]), _p(out["normal"]), _p(out["centroid"]), _p(out["residual"]))
    out["n_accepted"] = int(n)
    return out


def knn(map_xyz, q_xyz, k=5):
    mp = f32(map_xyz).reshape(-1, 3); q = f32(q_xyz).reshape(-1, 3)
    m = q.shape[0]
    idx = np.zeros((m, k), np.int32); d2 = np.zeros((m, k), np.float32); found = np.zeros(m, np.int32)
    lib().orc_knn(_p(mp), C.c_size_t(mp.shape[0]), _p(q), C.c_size_t(m), int(k), _p(idx), _p(d2), _p(found))
    return idx, d2, found


def pko_scale(residuals, cfg=None):
    cfg = cfg or default_icp_cfg()
    r = np.ascontiguousarray(residuals, np.float64)
    ns = max(cfg.gmm_sample_size, 1)
    sample = np.zeros(max(ns, len(r)), np.float64); n_sample = C.c_int(0)
    K = cfg.gmm_components
    means = np.zeros(K); var = np.zeros(K); w = np.zeros(K); em = C.c_int(0)
    js = np.zeros(cfg.num_alpha_segments + 1)
    a = lib().orc_pko_scale(_p(r), C.c_size_t(len(r)), C.byref(cfg), _p(sample), C.byref(n_sample), _p(means), _p(var), _p(w), C.byref(em), _p(js))
    return dict(alpha=float(a), sample=sample[: n_sample.value].copy(), means=means, vars=var, weights=w, em_iters=em.value, js=js)


def shuffle_head(n, head=100):
    out = np.zeros(min(head, n), np.int32)
    lib().orc_shuffle_head(int(n), int(head), _p(out))
    return out


def svd3f(A):
    A = f32(A).reshape(9); U = np.zeros(9, np.float32); S = np.zeros(3, np.float32); V = np.zeros(9, np.float32)
    lib().orc_svd3f(_p(A), _p(U), _p(S), _p(V))
    return U.reshape(3, 3), S, V.reshape(3, 3)


def so3_normalize(R):
    R = f32(R).reshape(9); o = np.zeros(9, np.float32)
    lib().orc_so3_normalize(_p(R), _p(o))
    return o.reshape(3, 3)


def so3_exp(w):
    w = f32(w); o = np.zeros(9, np.float32)
    lib().orc_so3_exp(_p(w), _p(o))
    return o.reshape(3, 3)


def ldlt6_solve(H, b):
    H = f32(H).reshape(36); b = f32(b); x = np.zeros(6, np.float32)
    lib().orc_ldlt6_solve(_p(H), _p(b), _p(x))
    return x


def se3_mul(A, B):
    A = f32(A).reshape(16); B = f32(B).reshape(16); o = np.zeros(16, np.float32)
    lib().orc_se3_mul(_p(A), _p(B), _p(o))
    return o.reshape(4, 4)


def se3_inv(A):
    A = f32(A).reshape(16); o = np.zeros(16, np.float32)
    lib().orc_se3_inv(_p(A), _p(o))
    return o.reshape(4, 4)


def fit_plane(cents):
    c = f32(cents).reshape(-1, 3)
    mu = np.zeros(3, np.float32); n = np.zeros(3, np.float32); pl = C.c_float(0)
    lib().orc_fit_plane(_p(c), int(c.shape[0]), _p(mu), _p(n), C.byref(pl))
    return mu, n, pl.value


class Pipeline:
    """Estimator-lite over the oracle (orc_pipeline.hpp)."""

    def __init__(self, cfg=None):
        self.cfg = cfg or default_pipe_cfg()
        self.h = C.c_void_p(lib().orc_pipe_create(C.byref(self.cfg)))

    def __del__(self):
        if self.h:
            lib().orc_pipe_destroy(self.h)
            self.h = None

    def process(self, scan):
        """scan: (N,3) or (N,4) float32.  Returns dict(pose, keyframe, icp_ok, times_ms, n_features, n_corr, n_iters)."""
        s = f32(scan)
        stride = s.shape[1]
        pose = np.zeros(16, np.float32); flags = C.c_int(0); times = np.zeros(4)
        nf, nc, ni = C.c_int(0), C.c_int(0), C.c_int(0)
        ok = lib().orc_pipe_process(self.h, _p(s), C.c_size_t(s.shape[0]), C.c_size_t(stride), _p(pose), C.byref(flags), _p(times),
                                    C.byref(nf), C.byref(nc), C.byref(ni))
        return dict(ok=bool(ok), pose=pose.reshape(4, 4).copy(), keyframe=bool(flags.value & 1), icp_ok=bool(flags.value & 2),
                    times_ms=times, n_features=nf.value, n_corr=nc.value, n_iters=ni.value)

    def map(self):
        return VoxelMap(handle=lib().orc_pipe_map(self.h))

    def features(self):
        n = lib().orc_pipe_features(self.h, None, C.c_size_t(0))
        out = np.zeros((n, 3), np.float32)
        lib().orc_pipe_features(self.h, _p(out), C.c_size_t(n))
        return out


def dense_order(ops):
    """Iteration order of the oracle's restated dense map under (op, key) pairs (op 0 = operator[], 1 = erase)."""
    ops = np.ascontiguousarray(ops, np.int64)
    out = np.zeros(len(ops) + 1, np.uint64)
    lib().orc_dense_order.restype = C.c_size_t
    n = lib().orc_dense_order(_p(ops), C.c_size_t(len(ops)), _p(out))
    return out[:n].copy()


# ---- the REAL reference pieces (oracle/_ref) -------------------------------------------------------
def ref_pko_available():
    return os.path.exists(REF_PKO)


def ref_pko_scale(residuals, cfg=None):
    cfg = cfg or default_icp_cfg()
    L = C.CDLL(REF_PKO)
    L.ref_pko_scale.restype = C.c_double
    L.ref_pko_scale.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int]
    r = np.ascontiguousarray(residuals, np.float64)
    return float(L.ref_pko_scale(_p(r), len(r), cfg.use_adaptive_m_estimator, cfg.loss_type, cfg.min_scale_factor, cfg.max_scale_factor,
                                 cfg.num_alpha_segments, cfg.truncated_threshold, cfg.gmm_components, cfg.gmm_sample_size, cfg.pko_kernel_type))


def ref_cont_available():
    return os.path.exists(REF_CONT)


def ref_dense_order(ops):
    """ops: (n,2) int64 (op, key); returns the iteration order of the real ankerl::unordered_dense::map."""
    L = C.CDLL(REF_CONT)
    L.ref_dense_order.restype = C.c_size_t
    ops = np.ascontiguousarray(ops, np.int64)
    out = np.zeros(len(ops) + 1, np.uint64)
    n = L.ref_dense_order(_p(ops), C.c_size_t(len(ops)), _p(out))
    return out[:n].copy()


def ref_knn(map_xyz, q_xyz, k=5):
    L = C.CDLL(REF_CONT)
    mp = f32(map_xyz).reshape(-1, 3); q = f32(q_xyz).reshape(-1, 3)
    m = q.shape[0]
    idx = np.zeros((m, k), np.int32); d2 = np.zeros((m, k), np.float32); found = np.zeros(m, np.int32)
    L.ref_knn(_p(mp), C.c_size_t(mp.shape[0]), _p(q), C.c_size_t(m), int(k), _p(idx), _p(d2), _p(found))
    return idx, d2, found
