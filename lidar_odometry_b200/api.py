"""Host-side mirror of the reference's hot-path classes over the C ABI (``include/b2lo.h``).

Same names, argument meaning and error behaviour as

* ``lidar_slam::map::FastVoxelFilter``                              /root/reference/src/database/VoxelMap.h:53-143
* ``lidar_slam::map::VoxelMap``                                     src/database/VoxelMap.h:188-332
* ``lidar_slam::optimization::ICPConfig`` / ``OptimizationStats`` /
  ``IterativeClosestPointOptimizer``                                src/optimization/IterativeClosestPointOptimizer.h:55-225
* ``lidar_slam::optimization::AdaptiveMEstimator`` (config carrier) src/optimization/AdaptiveMEstimator.h:28-143

so the parity tests read like tests of the reference.  Everything computes on the GPU through
``libb2lo.so``; numpy arrays are only the host buffers the reference's ``PointCloud`` would be.
The C++ drop-in with the identical class names lives in ``lidar_odometry_b200/shim/``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import capi
from .capi import B2LO_OK, B2LO_S_EMPTY, B2LO_S_INSUFFICIENT, IcpCfg, IcpStats, OdomCfg, OdomResult, RecordFmt, check

__all__ = ["PointShardedICP", "Context", "default_context", "FastVoxelFilter", "FastVoxelGrid", "VoxelGrid", "VoxelMap", "ICPConfig", "AdaptiveMEstimatorConfig",
           "AdaptiveMEstimator", "OptimizationStats", "IterativeClosestPointOptimizer", "Odometry", "OdometryBatch", "SE3", "RecordFormat", "kitti_record_format", "parse_ply_header", "load_ply_point_cloud", "load_kitti_binary"]


def _f32(a):
    if type(a) is np.ndarray and a.dtype == np.float32 and a.flags.c_contiguous:
        return a
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)   # the caller holds `a` for the duration of the call


def _cloud(a):
    """(N,3) or (N,4) float32 host cloud -> (array, n, stride_floats)."""
    a = _f32(a)
    if a.ndim != 2 or a.shape[1] < 3:
        raise ValueError("point cloud must be (N, >=3) float32")
    return a, a.shape[0], a.shape[1]


# ---- scan ingest: dataset records are the K1 input ---------------------------------------------------------------
@dataclass
class RecordFormat:
    """Where the three f32 coordinates sit in a fixed-size record of a scan file image (b2lo_record_fmt)."""
    record_bytes: int
    off_x: int
    off_y: int
    off_z: int

    def c(self) -> RecordFmt:
        return RecordFmt(self.record_bytes, self.off_x, self.off_y, self.off_z)


def kitti_record_format() -> RecordFormat:
    """KITTI .bin: x, y, z, intensity float32 (util::load_kitti_binary, src/util/PointCloudUtils.cpp:40-58)."""
    f = RecordFmt()
    capi.lib().b2lo_kitti_record_fmt(C.byref(f))
    return RecordFormat(f.record_bytes, f.off_x, f.off_y, f.off_z)


def _bytes(image):
    """file image (bytes / bytearray / uint8 array) -> contiguous uint8 array (no copy when it already is one)."""
    if isinstance(image, np.ndarray):
        a = image if image.dtype == np.uint8 else image.view(np.uint8)
        return np.ascontiguousarray(a).reshape(-1)
    return np.frombuffer(image, dtype=np.uint8)


def parse_ply_header(image):
    """PLYPlayer::parse_ply_header (app/player/ply_player.cpp:373-461) on a file image.  Returns a dict
    (fmt, vertex_count, data_offset, is_binary, n_records) or None for the files the reference rejects."""
    a = _bytes(image)
    f, vc, off, isb, nr = RecordFmt(), C.c_size_t(), C.c_size_t(), C.c_int(), C.c_size_t()
    rc = capi.lib().b2lo_ply_parse_header(_p(a), a.size, C.byref(f), C.byref(vc), C.byref(off), C.byref(isb), C.byref(nr))
    if rc != B2LO_OK:
        return None
    return dict(fmt=RecordFormat(f.record_bytes, f.off_x, f.off_y, f.off_z), vertex_count=vc.value, data_offset=off.value,
                is_binary=bool(isb.value), n_records=nr.value)


def load_ply_point_cloud(image):
    """PLYPlayer::load_ply_point_cloud (ply_player.cpp:267-371) as a HOST convenience: (N,3) float32, empty when the file is rejected.
    The hot path does not need it for binary files: hand the body to FastVoxelFilter.filter_records / Odometry.set_record_format."""
    a = _bytes(image)
    h = parse_ply_header(a)
    if h is None:
        return np.zeros((0, 3), np.float32)
    if h["is_binary"]:
        f, n = h["fmt"], h["n_records"]
        body = a[h["data_offset"]: h["data_offset"] + n * f.record_bytes].reshape(n, f.record_bytes)
        cols = [np.ascontiguousarray(body[:, o:o + 4]).view(np.float32).reshape(-1) for o in (f.off_x, f.off_y, f.off_z)]
        return np.stack(cols, axis=1) if n else np.zeros((0, 3), np.float32)
    out = np.zeros((h["vertex_count"], 3), np.float32)
    m = C.c_size_t()
    check(capi.lib().b2lo_ply_read_ascii(_p(a), a.size, _p(out), out.shape[0], C.byref(m)))
    return out[: m.value].copy()


def load_kitti_binary(path):
    """util::load_kitti_binary: the file image as (N,4) float32 records; every consumer here takes it as it is (stride 4)."""
    raw = np.fromfile(path, dtype=np.float32)
    return raw[: (raw.size // 4) * 4].reshape(-1, 4)


class Context:
    """One CUDA device + stream (b2lo_ctx)."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        check(capi.lib().b2lo_ctx_create(int(device), C.byref(self.h)))
        self.device = device

    def close(self):
        if self.h:
            capi.lib().b2lo_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(capi.lib().b2lo_ctx_sync(self.h))

    @property
    def stream(self):
        return capi.lib().b2lo_ctx_stream(self.h)

    @property
    def launch_count(self):
        return int(capi.lib().b2lo_ctx_launch_count(self.h))

    def io_bytes(self):
        a, b = C.c_ulonglong(), C.c_ulonglong()
        check(capi.lib().b2lo_ctx_io_bytes(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def features(self):
        """The feature cloud left on the device by the last filter run, as (M,3) float32."""
        m = C.c_size_t()
        check(capi.lib().b2lo_ctx_features(self.h, None, 0, C.byref(m)))
        out = np.zeros((m.value, 3), np.float32)
        if m.value:
            check(capi.lib().b2lo_ctx_features(self.h, _p(out), m.value, C.byref(m)))
        return out[: m.value]


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


# ---- SE3 helpers (util::SE3, MathUtils.h:101-168) ---------------------------------------------------------
class SE3:
    """Row-major 4x4 float32 poses with the reference's SE3 semantics (every product re-projects the rotation)."""

    @staticmethod
    def mul(A, B):
        A, B, out = _f32(A).reshape(16), _f32(B).reshape(16), np.zeros(16, np.float32)
        capi.lib().b2lo_se3_mul(_p(A), _p(B), _p(out))
        return out.reshape(4, 4)

    @staticmethod
    def inv(A):
        A, out = _f32(A).reshape(16), np.zeros(16, np.float32)
        capi.lib().b2lo_se3_inv(_p(A), _p(out))
        return out.reshape(4, 4)

    @staticmethod
    def from_rt(T):
        T, out = _f32(T).reshape(16), np.zeros(16, np.float32)
        capi.lib().b2lo_se3_from_rt(_p(T), _p(out))
        return out.reshape(4, 4)

    @staticmethod
    def log_so3(T):
        T, w = _f32(T).reshape(16), np.zeros(3, np.float32)
        capi.lib().b2lo_so3_log(_p(T), _p(w))
        return w

    @staticmethod
    def exp_so3(w):
        w, R = _f32(w).reshape(3), np.zeros(9, np.float32)
        capi.lib().b2lo_so3_exp(_p(w), _p(R))
        return R.reshape(3, 3)


# ---- FastVoxelFilter --------------------------------------------------------------------------------------
class FastVoxelFilter:
    def __init__(self, voxel_size=0.5, ctx=None):
        self.ctx = ctx or default_context()
        self.m_voxel_size = float(voxel_size)
        self._count = 0
        self.last_keys = None

    def setVoxelSize(self, voxel_size):
        self.m_voxel_size = float(voxel_size)

    def getVoxelSize(self):
        return self.m_voxel_size

    def filter(self, input, stride=1, want_keys=False):
        """Returns the downsampled cloud (M,3) in first-seen voxel order (VoxelMap.h:73-104)."""
        a, n, sf = _cloud(input)
        if n == 0:  # input.empty(): output.clear(); return
            self._count = 0
            self.last_keys = np.zeros(0, np.uint64)
            return np.zeros((0, 3), np.float32)
        ns = (n + stride - 1) // stride
        out = np.zeros((ns, 3), np.float32)
        keys = np.zeros(ns, np.uint64) if want_keys else None
        m = C.c_size_t()
        check(capi.lib().b2lo_filter(self.ctx.h, _p(a), n, sf, int(stride), C.c_float(self.m_voxel_size), _p(out), _p(keys), C.byref(m)))
        self._count = m.value
        self.last_keys = None if keys is None else keys[: m.value].copy()
        return out[: m.value].copy()

    def filter_records(self, image, fmt: RecordFormat, n_records=None, stride=1, want_keys=False, offset=0):
        """filter() over a record stream (KITTI .bin image, binary PLY body) read in place by K1 (b2lo_filter_records)."""
        a = _bytes(image)[offset:]
        if n_records is None:
            n_records = a.size // fmt.record_bytes
        if n_records * fmt.record_bytes > a.size:
            raise ValueError("record stream shorter than n_records")
        if n_records == 0:
            self._count = 0
            self.last_keys = np.zeros(0, np.uint64)
            return np.zeros((0, 3), np.float32)
        ns = (n_records + stride - 1) // stride
        out = np.zeros((ns, 3), np.float32)
        keys = np.zeros(ns, np.uint64) if want_keys else None
        m = C.c_size_t()
        f = fmt.c()
        check(capi.lib().b2lo_filter_records(self.ctx.h, _p(a), n_records, C.byref(f), int(stride), C.c_float(self.m_voxel_size), _p(out), _p(keys), C.byref(m)))
        self._count = m.value
        self.last_keys = None if keys is None else keys[: m.value].copy()
        return out[: m.value].copy()

    def getVoxelCount(self):
        return self._count


FastVoxelGrid = FastVoxelFilter  # VoxelMap.h:143


class VoxelGrid:
    """util::VoxelGrid (src/util/PointCloudUtils.h:462-557): setLeafSize / setInputCloud / filter, as save_map_to_ply uses it."""

    def __init__(self, ctx=None):
        self.ctx = ctx or default_context()
        self.leaf_size_ = 0.01
        self.input_cloud_ = None

    def setLeafSize(self, size):
        self.leaf_size_ = float(size)

    def setInputCloud(self, cloud):
        self.input_cloud_ = cloud

    def filter(self):
        """Returns the (M,3) centroids in the reference's std::map order."""
        if self.input_cloud_ is None or len(self.input_cloud_) == 0 or self.leaf_size_ <= 0:
            return np.zeros((0, 3), np.float32)
        a, n, sf = _cloud(self.input_cloud_)
        out = np.zeros((n, 3), np.float32)
        m = C.c_size_t()
        check(capi.lib().b2lo_voxel_grid_filter(self.ctx.h, _p(a), n, sf, C.c_float(self.leaf_size_), _p(out), n, C.byref(m)))
        return out[: m.value].copy()


# ---- VoxelMap ---------------------------------------------------------------------------------------------------
class VoxelMap:
    def __init__(self, voxel_size=0.5, ctx=None, capacity_hint=1 << 16):
        self.ctx = ctx or default_context()
        self.m_voxel_size = float(voxel_size)
        self.m_hierarchy_factor = 3
        self.m_planarity_threshold = 0.1   # VoxelMap.h:327
        self.m_compute_surfels = True      # VoxelMap.h:328
        self.m_max_hit_count = 10          # dead field in the reference (VoxelMap.h:296), kept for the setter
        self.m_init_hit_count = 1
        self._cap = int(capacity_hint)
        self.h = C.c_void_p()
        self._create()

    def _create(self):
        if self.h:
            capi.lib().b2lo_map_destroy(self.h)
            self.h = C.c_void_p()
        check(capi.lib().b2lo_map_create(self.ctx.h, C.c_float(self.m_voxel_size), self.m_hierarchy_factor, C.c_float(self.m_planarity_threshold),
                                         int(self.m_compute_surfels), self._cap, C.byref(self.h)))

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                capi.lib().b2lo_map_destroy(self.h)
        except Exception:
            pass

    # setters (VoxelMap.cpp:27-48, VoxelMap.h:197-209)
    def SetVoxelSize(self, size):
        if size <= 0:
            raise ValueError("Voxel size must be positive")  # std::invalid_argument
        if abs(self.m_voxel_size - size) > 1e-6:
            self.m_voxel_size = float(size)
            self._create()  # Clear() + new cell size

    def SetMaxHitCount(self, c):
        self.m_max_hit_count = int(c)

    def SetInitHitCount(self, c):
        self.m_init_hit_count = int(c)

    def SetHierarchyFactor(self, f):
        if f <= 0 or f % 2 == 0:
            return  # LOG_ERROR + ignored
        if f != 3:
            raise NotImplementedError("this build specialises the 3x3x3 hierarchy the reference uses (Estimator.cpp:79)")
        # unchanged factor: nothing to clear

    def SetPlanarityThreshold(self, t):
        self.m_planarity_threshold = float(t)
        check(capi.lib().b2lo_map_set_planarity_threshold(self.h, C.c_float(t)))

    def SetComputeSurfels(self, on):
        self.m_compute_surfels = bool(on)
        check(capi.lib().b2lo_map_set_compute_surfels(self.h, int(bool(on))))

    def GetVoxelSize(self):
        return self.m_voxel_size

    def GetHierarchyFactor(self):
        return self.m_hierarchy_factor

    def GetComputeSurfels(self):
        return self.m_compute_surfels

    def _counts(self, surfels=False):
        a, b, c = C.c_size_t(), C.c_size_t(), C.c_size_t()
        check(capi.lib().b2lo_map_counts(self.h, C.byref(a), C.byref(b), C.byref(c) if surfels else None))
        return a.value, b.value, c.value

    def GetVoxelCount(self):
        return self._counts()[0]

    def GetL1VoxelCount(self):
        return self._counts()[1]

    def GetSurfelCount(self):
        return self._counts(True)[2]

    def empty(self):
        return self._counts()[0] == 0

    def Clear(self):
        check(capi.lib().b2lo_map_clear(self.h))

    def UpdateVoxelMap(self, new_cloud, sensor_position, max_distance, is_keyframe=True):
        if new_cloud is None or len(new_cloud) == 0:
            return  # VoxelMap.cpp:134-136
        if not is_keyframe:
            return  # :138-140
        a, n, sf = _cloud(new_cloud)
        s = np.ascontiguousarray(sensor_position, np.float64).reshape(3)
        check(capi.lib().b2lo_map_update(self.h, _p(a), n, sf, _p(s), C.c_double(max_distance)))

    def ApplyTransformAndRehash(self, transform):
        T = _f32(transform).reshape(16)
        check(capi.lib().b2lo_map_transform_rehash(self.h, _p(T)))

    def GetSurfelAtPoint(self, point):
        p, n, c = _f32(point).reshape(3), np.zeros(3, np.float32), np.zeros(3, np.float32)
        rc = check(capi.lib().b2lo_map_lookup(self.h, _p(p), _p(n), _p(c)))
        return rc == 1, n, c

    def GetPointCloud(self):
        """L0 centroids in the reference's dense (insertion / swap-erase) order."""
        return self.export_l0()[0]

    def export_l0(self):
        n0 = self.GetVoxelCount()
        xyz = np.zeros((n0, 3), np.float32); keys = np.zeros((n0, 3), np.int32); cnt = np.zeros(n0, np.int32)
        n = C.c_size_t()
        check(capi.lib().b2lo_map_export_l0(self.h, _p(xyz), _p(keys), _p(cnt), n0, C.byref(n)))
        return xyz[: n.value], keys[: n.value], cnt[: n.value]

    def export_l1(self):
        n1 = self.GetL1VoxelCount()
        d = dict(keys=np.zeros((n1, 3), np.int32), nchild=np.zeros(n1, np.int32), children=np.zeros((n1, 27, 3), np.int32),
                 has_surfel=np.zeros(n1, np.int32), normal=np.zeros((n1, 3), np.float32), centroid=np.zeros((n1, 3), np.float32),
                 planarity=np.zeros(n1, np.float32), last_child_count=np.zeros(n1, np.int32))
        n = C.c_size_t()
        check(capi.lib().b2lo_map_export_l1(self.h, _p(d["keys"]), _p(d["nchild"]), _p(d["children"]), _p(d["has_surfel"]), _p(d["normal"]),
                                            _p(d["centroid"]), _p(d["planarity"]), _p(d["last_child_count"]), n1, C.byref(n)))
        return {k: v[: n.value] for k, v in d.items()}

    def GetL1Surfels(self):
        """List of (centroid, normal, planarity_score) — VoxelMap.cpp:405-418."""
        n1 = self.GetL1VoxelCount()
        c = np.zeros((n1, 3), np.float32); nr = np.zeros((n1, 3), np.float32); pl = np.zeros(n1, np.float32)
        n = C.c_size_t()
        check(capi.lib().b2lo_map_export_surfels(self.h, _p(c), _p(nr), _p(pl), None, n1, C.byref(n)))
        return [(c[i], nr[i], float(pl[i])) for i in range(n.value)]

    def RebuildKdTree(self):
        check(capi.lib().b2lo_map_rebuild_knn(self.h))

    def HasKdTree(self):
        return bool(capi.lib().b2lo_map_has_knn(self.h))

    def GetKdTree(self):
        return self if self.HasKdTree() else None  # the L0 hash itself is the search structure


# ---- ICP ---------------------------------------------------------------------------------------------------------
@dataclass
class ICPConfig:  # IterativeClosestPointOptimizer.h:55-76, defaults as wired by Estimator.cpp:62-70
    max_iterations: int = 4
    translation_tolerance: float = 0.005
    rotation_tolerance: float = 0.005
    max_correspondence_distance: float = 1.0
    min_correspondence_points: int = 10
    use_robust_loss: bool = True
    robust_loss_delta: float = 0.1
    use_surfel_correspondence: bool = True


@dataclass
class AdaptiveMEstimatorConfig:  # AdaptiveMEstimator.h:28-45 with config/kitti.yaml:42-51
    use_adaptive_m_estimator: bool = True
    loss_type: str = "huber"
    min_scale_factor: float = 0.1
    max_scale_factor: float = 10.0
    num_alpha_segments: int = 100
    truncated_threshold: float = 10.0
    gmm_components: int = 3
    gmm_sample_size: int = 100
    pko_kernel_type: str = "huber"


class AdaptiveMEstimator:
    """Config carrier: the PKO scale selection itself runs inside the device-resident Gauss-Newton loop."""

    def __init__(self, *args, **kw):
        self.config = args[0] if args and isinstance(args[0], AdaptiveMEstimatorConfig) else AdaptiveMEstimatorConfig(*args, **kw)

    def get_config(self):
        return self.config

    def reset(self):
        pass


@dataclass
class OptimizationStats:  # IterativeClosestPointOptimizer.h:203-210 (+ per-iteration taps)
    num_iterations: int = 0
    num_correspondences: int = 0
    initial_cost: float = 0.0
    final_cost: float = 0.0
    converged: bool = False
    optimization_time_ms: float = 0.0
    iterations: list = field(default_factory=list)


_KERNELS = {"huber": 0, "cauchy": 1}


def _icp_cfg(cfg: ICPConfig, ame: AdaptiveMEstimatorConfig | None) -> IcpCfg:
    c = IcpCfg()
    capi.lib().b2lo_default_icp_cfg(C.byref(c))
    c.max_iterations = cfg.max_iterations
    c.translation_tolerance = cfg.translation_tolerance
    c.rotation_tolerance = cfg.rotation_tolerance
    c.max_correspondence_distance = cfg.max_correspondence_distance
    c.min_correspondence_points = cfg.min_correspondence_points
    c.use_robust_loss = int(cfg.use_robust_loss)
    c.robust_loss_delta = cfg.robust_loss_delta
    c.use_surfel_correspondence = int(cfg.use_surfel_correspondence)
    if ame is None:  # no estimator attached: fixed delta, huber weights (ICP.cpp:319,394)
        c.use_adaptive_m_estimator = 0
        c.loss_type = 0
    else:
        if ame.pko_kernel_type not in _KERNELS:
            raise NotImplementedError(f"PKO kernel '{ame.pko_kernel_type}' (the reference configs use 'huber'; 'cauchy' also available)")
        c.use_adaptive_m_estimator = int(ame.use_adaptive_m_estimator)
        c.loss_type = 1 if ame.loss_type == "cauchy" else 0
        c.min_scale_factor = ame.min_scale_factor
        c.max_scale_factor = ame.max_scale_factor
        c.num_alpha_segments = ame.num_alpha_segments
        c.truncated_threshold = ame.truncated_threshold
        c.gmm_components = ame.gmm_components
        c.gmm_sample_size = ame.gmm_sample_size
        c.pko_kernel_type = _KERNELS[ame.pko_kernel_type]
    return c


def _trace(st: IcpStats):
    out = []
    for i in range(min(st.num_iterations, capi.B2LO_MAX_ITERS)):
        t = st.it[i]
        out.append(dict(n_corr=t.n_corr, scale=t.scale, delta=t.delta, H=np.array(t.H).reshape(6, 6), g=np.array(t.g), cost=t.cost,
                        dx=np.array(t.dx, np.float32), T_in=np.array(t.T_in, np.float32).reshape(4, 4),
                        T_out=np.array(t.T_out, np.float32).reshape(4, 4), em_iters=t.em_iters, kmeans_iters=t.kmeans_iters))
    return out


class IterativeClosestPointOptimizer:
    def __init__(self, config: ICPConfig | None = None, adaptive_estimator: AdaptiveMEstimator | None = None):
        self.m_config = config or ICPConfig()
        self.m_adaptive_estimator = adaptive_estimator
        self.m_last_stats = OptimizationStats()

    def update_config(self, config):
        self.m_config = config

    def get_config(self):
        return self.m_config

    def get_last_stats(self):
        return self.m_last_stats

    def optimize(self, voxel_map: VoxelMap, curr_frame, initial_transform):
        """Scan-to-map ICP (ICP.cpp:255-463).  ``curr_frame`` is a sensor-frame cloud (N,3|4) or an object with
        ``get_feature_cloud()`` / ``set_pose()`` like database::LidarFrame.  Returns ``(success, optimized_transform)``;
        on failure the transform is the initial one, as in the reference."""
        frame = curr_frame
        cloud = frame.get_feature_cloud() if hasattr(frame, "get_feature_cloud") else frame
        T0 = _f32(initial_transform).reshape(16)
        Tout = np.zeros(16, np.float32)
        self.m_last_stats = OptimizationStats()
        if voxel_map is None or cloud is None or len(cloud) == 0:
            return False, T0.reshape(4, 4).copy()
        a, n, sf = _cloud(cloud)
        ame = self.m_adaptive_estimator.get_config() if self.m_adaptive_estimator else None
        cfg = _icp_cfg(self.m_config, ame)
        st = IcpStats()
        rc = check(capi.lib().b2lo_icp_optimize(voxel_map.h, _p(a), n, sf, _p(T0), C.byref(cfg), _p(Tout), C.byref(st)))
        ok = rc == B2LO_OK
        self.m_last_stats = OptimizationStats(num_iterations=st.num_iterations, num_correspondences=st.num_correspondences,
                                              initial_cost=st.initial_cost, final_cost=st.final_cost, converged=ok,
                                              optimization_time_ms=st.device_ms, iterations=_trace(st))
        T = Tout.reshape(4, 4).copy()
        if hasattr(frame, "set_pose"):
            frame.set_pose(T)
        return ok, T

    def optimize_loop(self, curr_keyframe, matched_keyframe, ctx=None):
        """Loop-closure ICP between two keyframes (optimize_loop, ICP.cpp:40-251).  Each keyframe is ``(local_feature_cloud, pose)``
        or an object with ``get_feature_cloud()`` / ``get_pose()`` like database::LidarFrame.  Returns
        ``(success, optimized_relative_transform, inlier_ratio)`` with the reference's meaning: relative = curr_pose^-1 * optimised
        curr pose, success only if the loop converged within 100 iterations and at least half of the points are inliers."""
        def parts(kf):
            if hasattr(kf, "get_feature_cloud"):
                return kf.get_feature_cloud(), kf.get_pose()
            return kf
        cc, Tc = parts(curr_keyframe)
        mc, Tm = parts(matched_keyframe)
        ctx = ctx or default_context()
        Tc, Tm = _f32(Tc).reshape(16), _f32(Tm).reshape(16)
        Trel = np.zeros(16, np.float32)
        ratio = C.c_float(0.0)
        st = IcpStats()
        ame = self.m_adaptive_estimator.get_config() if self.m_adaptive_estimator else None
        cfg = _icp_cfg(self.m_config, ame)
        self.m_last_stats = OptimizationStats()
        if cc is None or mc is None or len(cc) == 0 or len(mc) == 0:
            return False, np.eye(4, dtype=np.float32), 0.0
        a, n, sf = _cloud(cc)
        b, m, sg = _cloud(mc)
        rc = check(capi.lib().b2lo_icp_optimize_loop(ctx.h, _p(a), n, sf, _p(Tc), _p(b), m, sg, _p(Tm), C.byref(cfg), _p(Trel), C.byref(ratio), C.byref(st)))
        self.m_last_stats = OptimizationStats(num_iterations=st.num_iterations, num_correspondences=st.num_correspondences,
                                              initial_cost=st.initial_cost, final_cost=st.final_cost, converged=bool(st.converged),
                                              optimization_time_ms=st.device_ms, iterations=_trace(st))
        return rc == B2LO_OK, Trel.reshape(4, 4).copy(), float(ratio.value)

    # parity taps -----------------------------------------------------------------------------------------------
    def iterate(self, voxel_map: VoxelMap, cloud, pose, scale=0.0):
        """ONE Gauss-Newton iteration (the loop body ICP.cpp:280-448) entered at ``pose`` with the residual normalisation ``scale`` an
        earlier iteration fixed (``scale <= 0``: computed here, as iteration 0 does).  Returns ``(status_ok, T_out, trace_dict)``."""
        a, n, sf = _cloud(cloud)
        T0 = _f32(pose).reshape(16)
        Tout = np.zeros(16, np.float32)
        ame = self.m_adaptive_estimator.get_config() if self.m_adaptive_estimator else None
        cfg = _icp_cfg(self.m_config, ame)
        st = IcpStats()
        rc = check(capi.lib().b2lo_icp_iterate(voxel_map.h, _p(a), n, sf, _p(T0), C.c_double(scale), C.byref(cfg), _p(Tout), C.byref(st)))
        tr = _trace(st)
        return rc == B2LO_OK, Tout.reshape(4, 4).copy(), (tr[0] if tr else None)

    def find_correspondences(self, voxel_map: VoxelMap, cloud, pose):
        """Per-query view of find_correspondences (ICP.cpp:587-645) at a fixed pose."""
        a, m, sf = _cloud(cloud)
        T = _f32(pose).reshape(16)
        out = dict(state=np.zeros(m, np.int32), l1key=np.zeros((m, 3), np.int32), morton=np.zeros(m, np.uint64),
                   normal=np.zeros((m, 3), np.float32), centroid=np.zeros((m, 3), np.float32), residual=np.zeros(m, np.float64))
        na = C.c_size_t()
        check(capi.lib().b2lo_icp_correspondences(voxel_map.h, _p(a), m, sf, _p(T), C.c_double(self.m_config.max_correspondence_distance),
                                                  _p(out["state"]), _p(out["l1key"]), _p(out["morton"]), _p(out["normal"]), _p(out["centroid"]),
                                                  _p(out["residual"]), C.byref(na)))
        out["n_accepted"] = na.value
        return out

    def find_correspondences_kdtree(self, voxel_map: VoxelMap, cloud, pose):
        """Per-query view of find_correspondences_kdtree (ICP.cpp:647-767) at a fixed pose."""
        a, m, sf = _cloud(cloud)
        T = _f32(pose).reshape(16)
        out = dict(knn=np.zeros((m, 5), np.int32), d2=np.zeros((m, 5), np.float32), found=np.zeros(m, np.int32), state=np.zeros(m, np.int32),
                   normal=np.zeros((m, 3), np.float32), centroid=np.zeros((m, 3), np.float32), residual=np.zeros(m, np.float64))
        na, nsc = C.c_size_t(), C.c_size_t()
        check(capi.lib().b2lo_icp_correspondences_knn(voxel_map.h, _p(a), m, sf, _p(T), C.c_double(self.m_config.max_correspondence_distance),
                                                      _p(out["knn"]), _p(out["d2"]), _p(out["found"]), _p(out["state"]), _p(out["normal"]),
                                                      _p(out["centroid"]), _p(out["residual"]), C.byref(na), C.byref(nsc)))
        out["n_accepted"] = na.value
        out["n_scanned"] = nsc.value
        return out


# PointShardedICP (the point-sharded mode) lives in sharding.py and is re-exported at the bottom of this module


# ---- per-scan driver ---------------------------------------------------------------------------------
class Odometry:
    """process_frame of processing::Estimator restricted to the hot path (Estimator.cpp:116-233), scan kept on the device."""

    def __init__(self, ctx=None, mid360=False, cfg: OdomCfg | None = None):
        self.ctx = ctx or default_context()
        if cfg is None:
            cfg = OdomCfg()
            capi.lib().b2lo_default_odom_cfg(C.byref(cfg), int(mid360))
        self.cfg = cfg
        self.h = C.c_void_p()
        check(capi.lib().b2lo_odom_create(self.ctx.h, C.byref(cfg), C.byref(self.h)))

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                capi.lib().b2lo_odom_destroy(self.h)
        except Exception:
            pass

    def reset(self):
        check(capi.lib().b2lo_odom_reset(self.h))

    def graph_stats(self):
        a, b, c = C.c_longlong(), C.c_longlong(), C.c_longlong()
        check(capi.lib().b2lo_odom_graph_stats(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(replays=a.value, builds=b.value, kernels_per_replay=c.value)

    @staticmethod
    def _result(rc, r: OdomResult):
        return dict(ok=rc == B2LO_OK, pose=np.frombuffer(r.pose, dtype=np.float32).reshape(4, 4).copy(), keyframe=bool(r.keyframe), icp_ok=r.icp_status == B2LO_OK,
                    n_features=r.n_features, n_corr=r.n_corr, n_iters=r.n_iters, device_ms=r.device_ms, l0=r.l0, l1=r.l1)

    def lookahead(self, next_scan) -> bool:
        """Announce the scan the NEXT process() call will get (b2lo_odom_lookahead): its K1 runs beside this scan's registration.
        next_scan: a page-locked host array, or (device_ptr, n, stride_floats).  Returns False when the announcement was ignored
        (pageable host memory).  The caller keeps the buffer alive and unchanged until that next call returns."""
        if isinstance(next_scan, tuple):
            ptr, n, sf = next_scan
            rc = check(capi.lib().b2lo_odom_lookahead(self.h, C.c_void_p(ptr), n, sf, 1))
        else:
            a, n, sf = _cloud(next_scan)
            if a is not next_scan and not (isinstance(next_scan, np.ndarray) and np.shares_memory(a, next_scan)):
                return False   # _cloud had to copy (dtype / layout): the copy would not outlive this call
            rc = check(capi.lib().b2lo_odom_lookahead(self.h, _p(a), n, sf, 0))
        return rc == B2LO_OK

    def set_record_format(self, fmt: RecordFormat | None):
        """Scans arrive as byte-record streams from now on (b2lo_odom_set_record_fmt); None returns to float clouds."""
        self._fmt = fmt
        if fmt is None:
            check(capi.lib().b2lo_odom_set_record_fmt(self.h, None))
        else:
            f = fmt.c()
            check(capi.lib().b2lo_odom_set_record_fmt(self.h, C.byref(f)))

    def process_records(self, image, n_records=None, offset=0, lookahead=None):
        """process() of one file image in the format given to set_record_format (uint8 array / bytes; offset = first record)."""
        fmt = getattr(self, "_fmt", None)
        if fmt is None:
            raise ValueError("set_record_format first")
        a = _bytes(image)[offset:]
        if n_records is None:
            n_records = a.size // fmt.record_bytes
        if n_records * fmt.record_bytes > a.size:
            raise ValueError("record stream shorter than n_records")
        if lookahead is not None:   # (uint8 array, n_records) of the next image; must be page-locked to take effect
            la, ln = lookahead
            check(capi.lib().b2lo_odom_lookahead(self.h, _p(la), ln, 3, 0))
        r = OdomResult()
        rc = check(capi.lib().b2lo_odom_process(self.h, _p(a), n_records, 3, C.byref(r)))
        return self._result(rc, r)

    def process(self, scan, lookahead=None):
        a, n, sf = _cloud(scan)
        r = OdomResult()
        if lookahead is not None and not isinstance(lookahead, tuple):
            b, bn, bsf = _cloud(lookahead)
            if b is lookahead or (isinstance(lookahead, np.ndarray) and np.shares_memory(b, lookahead)):   # else the copy would not outlive the call
                rc = check(capi.lib().b2lo_odom_process_la(self.h, a.ctypes.data, n, sf, b.ctypes.data, bn, bsf, C.byref(r)))   # one boundary crossing
                return self._result(rc, r)
        elif lookahead is not None:
            self.lookahead(lookahead)
        rc = check(capi.lib().b2lo_odom_process(self.h, _p(a), n, sf, C.byref(r)))
        return self._result(rc, r)

    def process_dev(self, dev_ptr, n, stride_floats, lookahead=None):
        if lookahead is not None:
            self.lookahead(lookahead)
        r = OdomResult()
        rc = check(capi.lib().b2lo_odom_process_dev(self.h, C.c_void_p(dev_ptr), n, stride_floats, C.byref(r)))
        return self._result(rc, r)

    def map(self):
        m = VoxelMap.__new__(VoxelMap)
        m.ctx = self.ctx
        m.h = C.c_void_p(capi.lib().b2lo_odom_map(self.h))
        m.m_voxel_size = self.cfg.map_voxel_size
        m.m_hierarchy_factor = 3
        m.m_planarity_threshold = self.cfg.surfel_planarity_threshold
        m.m_compute_surfels = bool(self.cfg.icp.use_surfel_correspondence)
        m.__class__ = _BorrowedVoxelMap
        return m


class OdometryBatch:
    """Independent sequences sharing one GPU, driven from one host thread (b2lo_odom_process_batch_dev): one scan per sequence per call,
    every sequence on its own context / stream / map / CUDA graphs."""

    def __init__(self, n_sequences, device=0, mid360=False):
        self.ctxs = [Context(device) for _ in range(n_sequences)]
        self.odos = [Odometry(c, mid360=mid360) for c in self.ctxs]
        self._h = (C.c_void_p * n_sequences)(*[o.h for o in self.odos])
        self._ptr = (C.c_void_p * n_sequences)()
        self._n = (C.c_size_t * n_sequences)()
        self._nptr = (C.c_void_p * n_sequences)()
        self._nn = (C.c_size_t * n_sequences)()
        self._res = (OdomResult * n_sequences)()

    def process_dev(self, dev_ptrs, ns, stride_floats, next_ptrs=None, next_ns=None):
        """dev_ptrs / ns: device address and point count of this call's scan of every sequence; next_*: the scans of the next call."""
        k = len(self.odos)
        for i in range(k):
            self._ptr[i] = dev_ptrs[i]; self._n[i] = ns[i]
            self._nptr[i] = next_ptrs[i] if next_ptrs is not None else None
            self._nn[i] = next_ns[i] if next_ns is not None else 0
        rc = check(capi.lib().b2lo_odom_process_batch_dev(self._h, self._ptr, self._n, self._nptr if next_ptrs is not None else None,
                                                           self._nn if next_ptrs is not None else None, stride_floats, k, self._res))
        return rc

    def results(self):
        return [Odometry._result(B2LO_OK if r.n_features else B2LO_S_EMPTY, r) for r in self._res]


class _BorrowedVoxelMap(VoxelMap):
    def __del__(self):  # owned by the Odometry handle
        pass


class LockstepBatch:
    """S independent sequences on one GPU advancing one scan per call, every kernel of the scan started once per call for all of them
    (b2lo_lockstep_*): the throughput mode for batches of recorded sequences.  ``odometries``: Odometry objects, one context each."""

    def __init__(self, odometries):
        self.ods = list(odometries)
        arr = (C.c_void_p * len(self.ods))(*[o.h for o in self.ods])
        self.h = C.c_void_p()
        check(capi.lib().b2lo_lockstep_create(arr, len(self.ods), C.byref(self.h)))
        self._res = (OdomResult * len(self.ods))()
        self._ptrs = (C.c_void_p * len(self.ods))()
        self._ns = (C.c_size_t * len(self.ods))()

    def __del__(self):
        if getattr(self, "h", None):
            try:
                capi.lib().b2lo_lockstep_destroy(self.h)
            except Exception:
                pass
            self.h = None

    def process_dev(self, scans, stride_floats, raw=False):
        """scans: one (device pointer, point count) per sequence.  Returns (list of result dicts, CUDA-event ms of the whole step)."""
        for i, (ptr, n) in enumerate(scans):
            self._ptrs[i] = ptr
            self._ns[i] = n
        ms = C.c_float(0.0)
        check(capi.lib().b2lo_lockstep_process_dev(self.h, self._ptrs, self._ns, stride_floats, self._res, C.byref(ms)))
        if raw:          # the b2lo_odom_result array itself (valid until the next call): no per-sequence Python objects on the hot loop
            return self._res, float(ms.value)
        out = []
        for r in self._res:
            out.append(dict(pose=np.array(r.pose, np.float32).reshape(4, 4), keyframe=bool(r.keyframe), icp_status=r.icp_status, n_features=r.n_features,
                            n_corr=r.n_corr, n_iters=r.n_iters, l0=r.l0, l1=r.l1))
        return out, float(ms.value)

    def stats(self):
        a, b, c, d = C.c_longlong(), C.c_longlong(), C.c_longlong(), C.c_longlong()
        check(capi.lib().b2lo_lockstep_stats(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return dict(kernels_per_step=a.value, replays=b.value, builds=c.value, fallbacks=d.value)


from .sharding import PointShardedICP  # noqa: E402  (the point-sharded mode; defined there, part of this interface)
