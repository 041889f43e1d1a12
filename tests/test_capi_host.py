"""CPU-side checks of the product boundary: libb2lo.so loads, exports every symbol include/b2lo.h declares, refuses to
work without a GPU (no CPU fallback), and its HOST-callable numerics (the same inline code the kernels run:
SE3/SO3 algebra, 3x3 Jacobi SVD, pivoted LDLT, surfel PCA, Z-order hash) agree bit-for-bit with the oracle."""
import ctypes as C
import re

import numpy as np
import pytest

from lidar_odometry_b200 import capi


def _names():
    hdr = open(capi.HEADER).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(b2lo_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    names = _names()
    assert len(names) >= 40
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/b2lo.h but not exported by libb2lo.so"
    assert set(names) == set(capi.SIGNATURES), set(names) ^ set(capi.SIGNATURES)
    assert b"sm_100a" in L.b2lo_version()


def test_struct_layouts_match_the_compiled_header():
    out = (C.c_size_t * 5)()
    capi.lib().b2lo_struct_sizes(out)
    mirrors = [capi.IcpCfg, capi.IterTrace, capi.IcpStats, capi.OdomCfg, capi.OdomResult]
    assert [C.sizeof(m) for m in mirrors] == list(out)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = capi.lib().b2lo_ctx_create(0, C.byref(h))
    assert rc == capi.B2LO_E_CUDA and not h.value
    assert "no CPU fallback" in capi.last_error()
    from lidar_odometry_b200 import api
    with pytest.raises(capi.B2loError):
        api.Context(0)


def _rand_pose(rng):
    from lidar_odometry_b200 import synth
    T = synth.pose_matrix(*rng.uniform(-50, 50, 3), *rng.uniform(-1, 1, 3)).astype(np.float32)
    return T


def test_host_numerics_bit_exact_vs_oracle(orc):
    from lidar_odometry_b200.api import SE3, _p
    L = capi.lib()
    rng = np.random.default_rng(12)
    for _ in range(200):
        A, B = _rand_pose(rng), _rand_pose(rng)
        assert np.array_equal(SE3.mul(A, B).view(np.uint32), orc.se3_mul(A, B).view(np.uint32))
        assert np.array_equal(SE3.inv(A).view(np.uint32), orc.se3_inv(A).view(np.uint32))
        w = (rng.standard_normal(3) * rng.choice([1e-8, 1e-3, 0.2, 2.0])).astype(np.float32)
        assert np.array_equal(SE3.exp_so3(w).view(np.uint32), orc.so3_exp(w).view(np.uint32))
        M = rng.standard_normal((3, 3)).astype(np.float32) * np.float32(rng.choice([1e-6, 1.0, 1e4]))
        U = np.zeros(9, np.float32); S = np.zeros(3, np.float32); V = np.zeros(9, np.float32)
        Mc = np.ascontiguousarray(M.reshape(9))
        L.b2lo_svd3(_p(Mc), _p(U), _p(S), _p(V))
        Uo, So, Vo = orc.svd3f(M)
        assert np.array_equal(U.view(np.uint32), Uo.reshape(9).view(np.uint32)) and np.array_equal(S.view(np.uint32), So.view(np.uint32))
        assert np.array_equal(V.view(np.uint32), Vo.reshape(9).view(np.uint32))
        X = rng.standard_normal((30, 6))
        H = np.ascontiguousarray((X.T @ X).astype(np.float32).reshape(36)); b = rng.standard_normal(6).astype(np.float32)
        x = np.zeros(6, np.float32)
        L.b2lo_ldlt6_solve(_p(H), _p(b), _p(x))
        assert np.array_equal(x.view(np.uint32), orc.ldlt6_solve(H, b).view(np.uint32))
        n = int(rng.integers(3, 28))
        pts = np.ascontiguousarray((rng.standard_normal((n, 3)) * [2, 2, 0.05] + rng.uniform(-80, 80, 3)).astype(np.float32))
        mu = np.zeros(3, np.float32); nr = np.zeros(3, np.float32); pl = C.c_float()
        L.b2lo_fit_plane(_p(pts), n, _p(mu), _p(nr), C.byref(pl))
        mo, no, po = orc.fit_plane(pts)
        assert np.array_equal(mu.view(np.uint32), mo.view(np.uint32)) and np.array_equal(nr.view(np.uint32), no.view(np.uint32))
        assert np.float32(pl.value).view(np.uint32) == np.float32(po).view(np.uint32)
        k = rng.integers(-(1 << 20), 1 << 20, 3)
        assert L.b2lo_voxel_key_hash(int(k[0]), int(k[1]), int(k[2])) == orc.voxel_key_hash(*k)


def test_so3_log_and_reprojection(orc):
    from lidar_odometry_b200.api import SE3
    rng = np.random.default_rng(13)
    for _ in range(100):
        T = _rand_pose(rng)
        w = SE3.log_so3(T)
        assert np.allclose(SE3.exp_so3(w), T[:3, :3], atol=3e-6)
        noisy = T.copy(); noisy[:3, :3] += rng.normal(0, 1e-4, (3, 3)).astype(np.float32)
        P = SE3.from_rt(noisy)
        assert np.array_equal(P[:3, :3].view(np.uint32), orc.so3_normalize(noisy[:3, :3]).view(np.uint32))
        assert np.array_equal(P[:3, 3], noisy[:3, 3])


def test_default_configs_match_reference_wiring(orc):
    c = capi.IcpCfg(); capi.lib().b2lo_default_icp_cfg(C.byref(c))
    o = orc.default_icp_cfg()
    for f, _ in capi.IcpCfg._fields_:
        assert getattr(c, f) == getattr(o, f), f
    for mid in (0, 1):
        a = capi.OdomCfg(); capi.lib().b2lo_default_odom_cfg(C.byref(a), mid)
        b = orc.default_pipe_cfg(bool(mid))
        for f, _ in capi.OdomCfg._fields_[:-1]:
            assert getattr(a, f) == getattr(b, f), f
        assert a.icp.use_surfel_correspondence == b.icp.use_surfel_correspondence
    # Estimator.cpp:62-70: min_correspondence_points stays at the ICPConfig default 10 (the yaml's 50 is never forwarded)
    assert c.min_correspondence_points == 10 and c.max_iterations == 4
