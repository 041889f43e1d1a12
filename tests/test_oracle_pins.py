"""Pins the CPU oracle (oracle/) before anything trusts it.

The reference has no tests, golden vectors or fixtures of its own, so the pins are outputs of the
REAL reference pieces that compile in the build container (oracle/_ref: AdaptiveMEstimator.cpp, unordered_dense,
nanoflann), stored as fixtures in tests/golden/ by tests/golden/make_golden.py, plus known-answer vectors of the
oracle itself.  When oracle/_ref/*.so is present the live comparison runs as well.  Eigen-level numerics
(JacobiSVD / LDLT) stay "parity unpinned" (Eigen is absent from the image) and are checked for mathematical
correctness instead.
"""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _sets(z):
    out, o = [], 0
    for n in z["n"]:
        out.append(z["flat"][o:o + n]); o += n
    return out


def test_pko_matches_real_reference_fixture(orc):
    z = np.load(os.path.join(G, "ref_pko.npz"))
    for r, a in zip(_sets(z), z["alpha"]):
        assert orc.pko_scale(r)["alpha"] == a, f"n={len(r)}"
    cfg = orc.default_icp_cfg(); cfg.pko_kernel_type = 1
    for r, a in zip(_sets(z)[:6], z["alpha_cauchy"]):
        assert orc.pko_scale(r, cfg)["alpha"] == a


def test_pko_matches_real_reference_live(orc):
    if not orc.ref_pko_available():
        pytest.skip("oracle/_ref/libref_pko.so not built (no /root/reference)")
    rng = np.random.default_rng(1)
    for n in (12, 99, 100, 150, 3000, 8191):
        r = np.abs(rng.normal(0, 1 + (n % 5), n))
        assert orc.pko_scale(r)["alpha"] == orc.ref_pko_scale(r)


def test_dense_map_order_matches_unordered_dense(orc):
    z = np.load(os.path.join(G, "ref_dense.npz"))
    assert np.array_equal(orc.dense_order(z["ops"]), z["order"])
    if orc.ref_cont_available():
        rng = np.random.default_rng(2)
        ops = np.c_[rng.integers(0, 2, 3000), rng.integers(0, 200, 3000)].astype(np.int64)
        assert np.array_equal(orc.dense_order(ops), orc.ref_dense_order(ops))


def test_kdtree_matches_nanoflann(orc):
    z = np.load(os.path.join(G, "ref_knn.npz"))
    idx, d2, found = orc.knn(z["cloud"], z["q"], 5)
    assert np.array_equal(idx, z["idx"]) and np.array_equal(found, z["found"])
    assert np.array_equal(d2.view(np.uint32), z["d2"].view(np.uint32))
    idx_t, _, found_t = orc.knn(z["cloud"][:3], z["q"][:10], 5)   # fewer points than k
    assert np.array_equal(found_t, z["found_t"]) and np.array_equal(idx_t[:, :3], z["idx_t"][:, :3])
    # and against exact brute force
    c, q = z["cloud"].astype(np.float32), z["q"].astype(np.float32)
    d = ((q[:, None, :] - c[None, :, :]) ** 2)
    dd = (d[..., 0] + d[..., 1]) + d[..., 2]
    assert np.array_equal(np.sort(np.argsort(dd, axis=1, kind="stable")[:, :5], axis=1), np.sort(idx, axis=1))


def test_oracle_known_answers(orc):
    from lidar_odometry_b200 import synth
    z = np.load(os.path.join(G, "oracle_kat.npz"))
    scans, poses = synth.kitti_sequence(n_scans=3, seed=3, n_rings=32, n_az=400)
    feat, keys = orc.voxel_filter(scans[0][:, :3], 8, 0.5)
    assert np.array_equal(keys, z["keys"]) and np.array_equal(feat.view(np.uint32), z["feat"].view(np.uint32))
    pipe = orc.Pipeline()
    res = [pipe.process(s) for s in scans]
    assert np.array_equal(np.array([r["n_corr"] for r in res]), z["n_corr"])
    assert np.array_equal(np.stack([r["pose"] for r in res]).view(np.uint32), z["poses"].view(np.uint32))
    k0, c0, n0 = pipe.map().export_l0()
    assert np.array_equal(k0, z["l0_keys"]) and np.array_equal(n0, z["l0_cnt"]) and np.array_equal(c0.view(np.uint32), z["l0_cent"].view(np.uint32))
    for n in (1000, 65536, 4321):
        assert np.array_equal(orc.shuffle_head(n, 100), z[f"shuffle_head_{n}"])


def test_keys_and_hashes(orc):
    # FastVoxelFilter::computeMortonKey clamps (VoxelMap.h:124-135); VoxelKeyHash wraps (:166-183)
    assert orc.filter_morton_key(0.0, 0.0, 0.0, 0.5) == orc.voxel_key_hash(0, 0, 0)
    assert orc.filter_morton_key(-0.1, 0.0, 0.0, 0.5) == orc.voxel_key_hash(-1, 0, 0)
    assert orc.filter_morton_key(1e9, 0.0, 0.0, 0.5) == orc.voxel_key_hash((1 << 20) - 1, 0, 0)      # clamp high
    assert orc.filter_morton_key(-1e9, 0.0, 0.0, 0.5) == orc.voxel_key_hash(-(1 << 20), 0, 0)        # clamp low
    assert orc.voxel_key_hash(1 << 20, 0, 0) == orc.voxel_key_hash(-(1 << 20), 0, 0)                  # wrap
    # the 2^20 offset of every axis lands in bits 60/61/62 of the interleave; x is the lowest bit of each triple
    base = orc.voxel_key_hash(0, 0, 0)
    assert base == (1 << 60) | (1 << 61) | (1 << 62)
    assert orc.voxel_key_hash(1, 0, 0) == base + 1 and orc.voxel_key_hash(0, 1, 0) == base + 2 and orc.voxel_key_hash(0, 0, 1) == base + 4
    assert orc.voxel_key_hash(-1, 0, 0) == (base & ~(1 << 60)) | sum(1 << (3 * b) for b in range(20))
    # parents: floor division for negatives (VoxelMap.cpp:60-67)
    assert list(orc.parent_key([-1, -3, -4])) == [-1, -1, -2] and list(orc.parent_key([0, 2, 3])) == [0, 0, 1]
    # level-1 key uses a float division by voxel*3 (:50-58)
    assert list(orc.point_to_key([1.49, -0.01, 3.0], 0.5, 3, 1)) == [0, -1, 2]


def test_eigen_restatement_is_mathematically_sound(orc):
    rng = np.random.default_rng(4)
    for _ in range(50):
        A = rng.standard_normal((3, 3)).astype(np.float32)
        U, S, V = orc.svd3f(A)
        assert np.allclose(U @ np.diag(S) @ V.T, A, atol=2e-6) and S[0] >= S[1] >= S[2] >= 0
        assert np.allclose(S, np.linalg.svd(A.astype(np.float64), compute_uv=False), atol=2e-6)
        R = orc.so3_normalize(A)
        assert np.allclose(R @ R.T, np.eye(3), atol=2e-6) and np.linalg.det(R.astype(np.float64)) > 0.999
        M = rng.standard_normal((20, 6))
        H = (M.T @ M).astype(np.float32); b = rng.standard_normal(6).astype(np.float32)
        x = orc.ldlt6_solve(H, b)
        assert np.allclose(x, np.linalg.solve(H.astype(np.float64), b), rtol=2e-3, atol=2e-4)
        w = (rng.standard_normal(3) * 0.3).astype(np.float32)
        Rw = orc.so3_exp(w)
        th = np.linalg.norm(w); K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]) / th
        assert np.allclose(Rw, np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K, atol=2e-6)
