"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/_ref/libref_core.so: the UNMODIFIED reference translation units (VoxelMap.cpp, LidarFrame.cpp,
MathUtils.cpp, PointCloudUtils.cpp, IterativeClosestPointOptimizer.cpp, AdaptiveMEstimator.cpp) compiled where they lie under
/root/reference against oracle/eigen_compat (see oracle/Makefile, oracle/src/ref_core_wrap.cpp).  The entry points mirror
oracle/orc.py so that tests/test_oracle_pins.py and tests/golden/make_golden.py run restatement and reference on the same
inputs.  Exists only where the library was built (this container, and the GPU box through the gpurun snapshot); never
imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import orc
from .orc import IcpCfg, IterTrace, _p, f32

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CORE = os.path.join(HERE, "_ref", "libref_core.so")

_lib = None


def available():
    return os.path.exists(REF_CORE)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(REF_CORE)
        L.ref_voxel_key_hash.restype = C.c_uint64
        L.ref_voxel_key_hash.argtypes = [C.c_int, C.c_int, C.c_int]
        L.ref_map_create.restype = C.c_void_p
        L.ref_map_create.argtypes = [C.c_float, C.c_int, C.c_float, C.c_int]
        for name in ("ref_map_point_cloud", "ref_map_surfels", "ref_icp_correspondence_list"):
            getattr(L, name).restype = C.c_size_t
        L.ref_pipe_create.restype = C.c_void_p
        L.ref_pipe_create.argtypes = [C.c_void_p]
        L.ref_pipe_destroy.argtypes = [C.c_void_p]
        L.ref_pipe_process.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_pipe_map.restype = C.c_void_p
        L.ref_pipe_map.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def voxel_key_hash(x, y, z):
    return int(lib().ref_voxel_key_hash(int(x), int(y), int(z)))


def point_to_key(p, voxel, factor=3, level=0):
    p = f32(p)
    k = np.zeros(3, np.int32)
    lib().ref_point_to_key(_p(p), C.c_float(voxel), int(factor), int(level), _p(k))
    return k


def parent_key(key, factor=3):
    key = np.ascontiguousarray(key, np.int32)
    out = np.zeros(3, np.int32)
    lib().ref_parent_key(_p(key), int(factor), _p(out))
    return out


def voxel_filter(xyz, stride, voxel):
    xyz = f32(xyz).reshape(-1, 3)
    n = xyz.shape[0]
    out = np.zeros((max(n, 1), 3), np.float32)
    m = C.c_size_t(0)
    lib().ref_filter(_p(xyz), C.c_size_t(n), int(stride), C.c_float(voxel), _p(out), C.byref(m))
    return out[: m.value].copy()


def voxel_grid_filter(xyz, leaf):
    xyz = f32(xyz).reshape(-1, 3)
    n = xyz.shape[0]
    out = np.zeros((max(n, 1), 3), np.float32)
    m = C.c_size_t(0)
    rc = lib().ref_voxel_grid_filter(_p(xyz), C.c_size_t(n), C.c_float(leaf), _p(out), C.c_size_t(max(n, 1)), C.byref(m))
    assert rc == 0
    return out[: m.value].copy()


def kitti_load_file(path, cap_points):
    out = np.zeros((max(cap_points, 1), 3), np.float32)
    n = C.c_size_t(0)
    rc = lib().ref_kitti_load_file(str(path).encode(), _p(out), C.c_size_t(max(cap_points, 1)), C.byref(n))
    assert rc == 0, rc
    return out[: n.value].copy()


def transform_point_cloud(xyz, T):
    xyz = f32(xyz).reshape(-1, 3)
    T = f32(T).reshape(16)
    out = np.zeros_like(xyz)
    lib().ref_transform_point_cloud(_p(xyz), C.c_size_t(xyz.shape[0]), _p(T), _p(out))
    return out


class VoxelMap:
    """lidar_slam::map::VoxelMap (the real class)."""

    def __init__(self, voxel=0.5, factor=3, planarity=0.1, compute_surfels=True):
        self.h = C.c_void_p(lib().ref_map_create(voxel, factor, planarity, int(compute_surfels)))

    def __del__(self):
        if getattr(self, "h", None) and getattr(self, "owned", True):
            lib().ref_map_destroy(self.h)
        self.h = None

    def clear(self):
        lib().ref_map_clear(self.h)

    def update(self, xyz, sensor, max_distance):
        xyz = f32(xyz).reshape(-1, 3)
        s = np.ascontiguousarray(sensor, np.float64)
        lib().ref_map_update(self.h, _p(xyz), C.c_size_t(xyz.shape[0]), _p(s), C.c_double(max_distance))

    def counts(self):
        a, b, c = C.c_size_t(), C.c_size_t(), C.c_size_t()
        lib().ref_map_counts(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def export_l0(self):
        n = self.counts()[0]
        keys = np.zeros((n, 3), np.int32); cent = np.zeros((n, 3), np.float32); cnt = np.zeros(n, np.int32)
        lib().ref_map_export_l0(self.h, _p(keys), _p(cent), _p(cnt))
        return keys, cent, cnt

    def point_cloud(self):
        n = self.counts()[0]
        out = np.zeros((max(n, 1), 3), np.float32)
        k = lib().ref_map_point_cloud(self.h, _p(out), C.c_size_t(max(n, 1)))
        return out[:k].copy()

    def export_l1(self):
        n = self.counts()[1]
        d = dict(keys=np.zeros((n, 3), np.int32), nchild=np.zeros(n, np.int32), children=np.zeros((n, 27, 3), np.int32),
                 has_surfel=np.zeros(n, np.int32), normal=np.zeros((n, 3), np.float32), centroid=np.zeros((n, 3), np.float32),
                 planarity=np.zeros(n, np.float32), last_child_count=np.zeros(n, np.int32))
        lib().ref_map_export_l1(self.h, _p(d["keys"]), _p(d["nchild"]), _p(d["children"]), _p(d["has_surfel"]), _p(d["normal"]),
                                _p(d["centroid"]), _p(d["planarity"]), _p(d["last_child_count"]))
        return d

    def surfels(self):
        n = self.counts()[1]
        c = np.zeros((max(n, 1), 3), np.float32); nn = np.zeros((max(n, 1), 3), np.float32); pl = np.zeros(max(n, 1), np.float32)
        k = lib().ref_map_surfels(self.h, _p(c), _p(nn), _p(pl), C.c_size_t(max(n, 1)))
        return c[:k].copy(), nn[:k].copy(), pl[:k].copy()

    def lookup(self, p):
        p = f32(p); n = np.zeros(3, np.float32); c = np.zeros(3, np.float32)
        ok = lib().ref_map_lookup(self.h, _p(p), _p(n), _p(c))
        return bool(ok), n, c

    def transform_rehash(self, T):
        T = f32(T).reshape(16)
        lib().ref_map_transform_rehash(self.h, _p(T))

    def rebuild_kdtree(self):
        lib().ref_map_rebuild_kdtree(self.h)


def correspondence_list(vmap, local_xyz, T, max_dist=1.0, kdtree=False):
    """find_correspondences / find_correspondences_kdtree (private members) -> the DualFrameCorrespondences lists."""
    local = f32(local_xyz).reshape(-1, 3)
    m = local.shape[0]
    T = f32(T).reshape(16)
    pl = np.zeros((max(m, 1), 3)); pc = np.zeros((max(m, 1), 3)); nl = np.zeros((max(m, 1), 3)); r = np.zeros(max(m, 1))
    n = lib().ref_icp_correspondence_list(vmap.h, _p(local), C.c_size_t(m), _p(T), C.c_double(max_dist), int(kdtree), _p(pl), _p(pc), _p(nl),
                                          _p(r), C.c_size_t(max(m, 1)))
    return dict(points_last=pl[:n].copy(), points_curr=pc[:n].copy(), normals_last=nl[:n].copy(), residuals=r[:n].copy())


def icp_optimize(vmap, local_xyz, T_init, cfg=None, trace_cap=64):
    """optimize (ICP.cpp:255-463) -> (ok, T_out, trace[H, g, dx per iteration from the LDLT hook], stats, frame pose)."""
    cfg = cfg or orc.default_icp_cfg()
    local = f32(local_xyz).reshape(-1, 3)
    T0 = f32(T_init).reshape(16)
    Tout = np.zeros(16, np.float32); Tframe = np.zeros(16, np.float32)
    tr = (IterTrace * trace_cap)()
    nt = C.c_int(0)
    stats = np.zeros(6)
    ok = lib().ref_icp_optimize(vmap.h, _p(local), C.c_size_t(local.shape[0]), _p(T0), C.byref(cfg), _p(Tout), tr, trace_cap, C.byref(nt),
                                _p(stats), _p(Tframe))
    trace = [dict(H=np.array(tr[i].H, np.float32).reshape(6, 6), g=np.array(tr[i].g, np.float32), dx=np.array(tr[i].dx, np.float32))
             for i in range(nt.value)]
    st = dict(num_correspondences=int(stats[0]), num_iterations=int(stats[1]), initial_cost=stats[2], final_cost=stats[3],
              converged=bool(stats[4]))
    return bool(ok), Tout.reshape(4, 4), trace, st, Tframe.reshape(4, 4)


def icp_optimize_loop(curr_xyz, T_curr, matched_xyz, T_matched, cfg=None, trace_cap=128):
    cfg = cfg or orc.default_icp_cfg()
    cur = f32(curr_xyz).reshape(-1, 3)
    mat = f32(matched_xyz).reshape(-1, 3)
    Tc, Tm = f32(T_curr).reshape(16), f32(T_matched).reshape(16)
    Trel = np.zeros(16, np.float32)
    ratio, iters, nt = C.c_float(0), C.c_int(0), C.c_int(0)
    tr = (IterTrace * trace_cap)()
    ok = lib().ref_icp_optimize_loop(_p(cur), C.c_size_t(cur.shape[0]), _p(Tc), _p(mat), C.c_size_t(mat.shape[0]), _p(Tm), C.byref(cfg),
                                     _p(Trel), C.byref(ratio), C.byref(iters), tr, trace_cap, C.byref(nt))
    trace = [dict(H=np.array(tr[i].H, np.float32).reshape(6, 6), g=np.array(tr[i].g, np.float32), dx=np.array(tr[i].dx, np.float32))
             for i in range(nt.value)]
    return bool(ok), Trel.reshape(4, 4), float(ratio.value), int(iters.value), trace


def so3_normalize(R):
    R = f32(R).reshape(9); o = np.zeros(9, np.float32)
    lib().ref_so3_normalize(_p(R), _p(o))
    return o.reshape(3, 3)


def so3_exp(w):
    w = f32(w); o = np.zeros(9, np.float32)
    lib().ref_so3_exp(_p(w), _p(o))
    return o.reshape(3, 3)


def so3_log(R):
    R = f32(R).reshape(9); o = np.zeros(3, np.float32)
    lib().ref_so3_log(_p(R), _p(o))
    return o


def se3_mul(A, B):
    A = f32(A).reshape(16); B = f32(B).reshape(16); o = np.zeros(16, np.float32)
    lib().ref_se3_mul(_p(A), _p(B), _p(o))
    return o.reshape(4, 4)


def se3_inv(A):
    A = f32(A).reshape(16); o = np.zeros(16, np.float32)
    lib().ref_se3_inv(_p(A), _p(o))
    return o.reshape(4, 4)


def svd3f(A):
    A = f32(A).reshape(9); U = np.zeros(9, np.float32); S = np.zeros(3, np.float32); V = np.zeros(9, np.float32)
    lib().ref_svd3f(_p(A), _p(U), _p(S), _p(V))
    return U.reshape(3, 3), S, V.reshape(3, 3)


def plane_normal_nx3(A):
    A = np.ascontiguousarray(A, np.float64).reshape(-1, 3)
    n = np.zeros(3)
    lib().ref_plane_normal_nx3(_p(A), int(A.shape[0]), _p(n))
    return n


class Pipeline:
    """The per-scan driver (Estimator.cpp:116-233 control flow) over the reference's OWN classes; same interface as orc.Pipeline."""

    def __init__(self, cfg=None):
        self.cfg = cfg or orc.default_pipe_cfg()
        self.h = C.c_void_p(lib().ref_pipe_create(C.byref(self.cfg)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_pipe_destroy(self.h)
            self.h = None

    def process(self, scan):
        s = f32(scan)
        pose = np.zeros(16, np.float32); flags = C.c_int(0); times = np.zeros(4)
        nf, nc, ni = C.c_int(0), C.c_int(0), C.c_int(0)
        ok = lib().ref_pipe_process(self.h, _p(s), s.shape[0], s.shape[1], _p(pose), C.byref(flags), _p(times), C.byref(nf), C.byref(nc), C.byref(ni))
        return dict(ok=bool(ok), pose=pose.reshape(4, 4).copy(), keyframe=bool(flags.value & 1), icp_ok=bool(flags.value & 2),
                    times_ms=times, n_features=nf.value, n_corr=nc.value, n_iters=ni.value)

    def map(self):
        m = VoxelMap.__new__(VoxelMap)
        m.h = C.c_void_p(lib().ref_pipe_map(self.h)); m.owned = False; m._pipe = self   # borrowed: the pipeline owns it
        return m


# ---- the reference's own PLY reader (oracle/_ref/libref_ply.so: the unmodified app/player/ply_player.cpp) ---------------------------
REF_PLY = os.path.join(HERE, "_ref", "libref_ply.so")
_ply_lib = None


def ply_available():
    return os.path.exists(REF_PLY)


def _ply():
    global _ply_lib
    if _ply_lib is None:
        L = C.CDLL(REF_PLY)
        L.ref_ply_load_file.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.ref_ply_parse_header.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _ply_lib = L
    return _ply_lib


def ply_load_file(path, cap=1 << 20):
    """PLYPlayer::load_ply_point_cloud(path) -> (N,3) f32 (empty when the reference returns nullptr / an empty cloud)."""
    out = np.zeros((cap, 3), np.float32)
    n = C.c_size_t(0)
    rc = _ply().ref_ply_load_file(os.fsencode(path), _p(out), cap, C.byref(n))
    if rc < 0:
        raise RuntimeError("ref_ply_load_file: buffer too small")
    return out[: n.value].copy() if rc == 1 else np.zeros((0, 3), np.float32)


def ply_parse_header(path):
    """PLYPlayer::parse_ply_header(path) -> dict(ok, vertex_count, is_binary, n_props, stride)."""
    vc, st = C.c_size_t(0), C.c_size_t(0)
    b, npr = C.c_int(0), C.c_int(0)
    ok = _ply().ref_ply_parse_header(os.fsencode(path), C.byref(vc), C.byref(b), C.byref(npr), C.byref(st))
    return dict(ok=bool(ok), vertex_count=vc.value, is_binary=bool(b.value), n_props=npr.value, stride=st.value)


# ---- the reference's own Estimator (oracle/_ref/libref_estimator.so: the unmodified src/processing/Estimator.cpp) --------------------
REF_EST = os.path.join(HERE, "_ref", "libref_estimator.so")
_est_lib = None


def estimator_available():
    return os.path.exists(REF_EST) and available()


def _est():
    global _est_lib
    if _est_lib is None:
        lib()                                      # libref_core.so first (the estimator library links it)
        L = C.CDLL(REF_EST)
        L.ref_est_create.restype = C.c_void_p
        L.ref_est_create.argtypes = [C.c_void_p]
        L.ref_est_destroy.argtypes = [C.c_void_p]
        L.ref_est_process.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_est_map.restype = C.c_void_p
        L.ref_est_map.argtypes = [C.c_void_p]
        _est_lib = L
    return _est_lib


REF_EST_GPU = os.path.join(HERE, "_ref", "libref_estimator_gpu.so")
_est_gpu_lib = None


def estimator_gpu_available():
    return os.path.exists(REF_EST_GPU)


def _est_gpu():
    """The same Estimator.cpp compiled against the drop-in shim and linked with libb2lo.so (needs a GPU to run)."""
    global _est_gpu_lib
    if _est_gpu_lib is None:
        L = C.CDLL(REF_EST_GPU)
        L.ref_est_create.restype = C.c_void_p
        L.ref_est_create.argtypes = [C.c_void_p]
        L.ref_est_destroy.argtypes = [C.c_void_p]
        L.ref_est_process.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_est_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _est_gpu_lib = L
    return _est_gpu_lib


class Estimator:
    """processing::Estimator itself (loop detection and pose-graph optimisation switched off): process() = process_frame on one scan.
    gpu=True: the build in which database/VoxelMap.h and optimization/IterativeClosestPointOptimizer.h are the drop-in shim
    (oracle/_ref/libref_estimator_gpu.so) - the reference's own driver running the CUDA engine."""

    def __init__(self, cfg=None, gpu=False):
        self.cfg = cfg or orc.default_pipe_cfg()
        self._L = _est_gpu() if gpu else _est()
        self.gpu = gpu
        self.h = C.c_void_p(self._L.ref_est_create(C.byref(self.cfg)))

    def __del__(self):
        if getattr(self, "h", None):
            self._L.ref_est_destroy(self.h)
            self.h = None

    def counts(self):
        a, b, c = C.c_size_t(), C.c_size_t(), C.c_size_t()
        self._L.ref_est_counts(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def last_process_ms(self):
        """Wall time of the last process_frame call alone (without the binding's cloud conversion); None with an older library."""
        f = getattr(self._L, "ref_est_last_process_ms", None)
        if f is None:
            return None
        f.restype = C.c_double
        f.argtypes = [C.c_void_p]
        return float(f(self.h))

    def process(self, scan):
        s = f32(scan)
        pose = np.zeros(16, np.float32); flags = C.c_int(0); nf = C.c_int(0)
        ok = self._L.ref_est_process(self.h, _p(s), s.shape[0], s.shape[1], _p(pose), C.byref(flags), C.byref(nf))
        return dict(ok=bool(ok), pose=pose.reshape(4, 4).copy(), keyframe=bool(flags.value & 1), n_features=nf.value)

    def map(self):
        m = VoxelMap.__new__(VoxelMap)
        assert not self.gpu, "the GPU build's map is the shim's class: use counts()"
        m.h = C.c_void_p(_est().ref_est_map(self.h)); m.owned = False; m._est = self   # borrowed: the estimator owns it
        return m
