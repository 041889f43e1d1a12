"""Generates tests/golden/*.npz.  Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

* ref_pko.npz    — outputs of the REAL reference AdaptiveMEstimator (oracle/_ref/libref_pko.so, compiled in place from
                   /root/reference/src/optimization/AdaptiveMEstimator.cpp): residual vectors -> alpha.
* ref_dense.npz  — iteration order of the REAL ankerl::unordered_dense::map under insert/erase sequences.
* ref_knn.npz    — kNN indices/distances of the REAL nanoflann kd-tree (leaf 10).
* ref_core.npz   — outputs of the REAL reference VoxelMap / FastVoxelFilter / IterativeClosestPointOptimizer::optimize
                   (oracle/_ref/libref_core.so: the unmodified reference sources compiled against oracle/eigen_compat) on a seeded
                   5-scan sequence: per-keyframe digests of the whole map state, the final map, H / g / dx of every Gauss-Newton
                   iteration, the poses.  tests/test_reference_core.py::golden_run is the single definition of that run.
* oracle_kat.npz — known-answer vectors of the oracle itself on seeded inputs (regression pin that travels to the GPU box).
Nothing here is read from /root/reference at test time; the fixtures are committed.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402
from lidar_odometry_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def residual_sets():
    rng = np.random.default_rng(2024)
    sets = []
    for n in (7, 10, 63, 100, 101, 512, 1000, 2047, 4096, 9973, 65535, 65536, 70001):
        kind = n % 3
        if kind == 0:
            r = np.abs(rng.normal(0, 1.0, n))
        elif kind == 1:
            r = np.abs(np.r_[rng.normal(0, 0.5, n - n // 5), rng.normal(4, 1.5, n // 5)])
            rng.shuffle(r)
        else:
            r = rng.gamma(1.5, 1.2, n)
        sets.append(r.astype(np.float64))
    return sets


def main():
    orc.build()
    assert orc.ref_pko_available() and orc.ref_cont_available(), "needs /root/reference to build oracle/_ref"
    # --- real reference PKO
    sets = residual_sets()
    alphas = np.array([orc.ref_pko_scale(r) for r in sets])
    cfg = orc.default_icp_cfg(); cfg.pko_kernel_type = 1
    alphas_cauchy = np.array([orc.ref_pko_scale(r, cfg) for r in sets[:6]])
    np.savez_compressed(os.path.join(OUT, "ref_pko.npz"), n=np.array([len(r) for r in sets]), flat=np.concatenate(sets), alpha=alphas,
                        alpha_cauchy=alphas_cauchy)
    # --- real unordered_dense ordering
    rng = np.random.default_rng(7)
    ops = []
    live = []
    for i in range(4000):
        if live and rng.uniform() < 0.35:
            k = live.pop(int(rng.integers(len(live))))
            ops.append((1, k))
        else:
            k = int(rng.integers(0, 1 << 40))
            if rng.uniform() < 0.2 and live:
                k = live[int(rng.integers(len(live)))]
            else:
                live.append(k)
            ops.append((0, k))
    ops = np.array(ops, np.int64)
    order = orc.ref_dense_order(ops)
    np.savez_compressed(os.path.join(OUT, "ref_dense.npz"), ops=ops, order=order)
    # --- real nanoflann
    rng = np.random.default_rng(11)
    cloud = (rng.uniform(-10, 10, (3000, 3)) * [1, 1, 0.1]).astype(np.float32)
    q = (rng.uniform(-12, 12, (400, 3)) * [1, 1, 0.2]).astype(np.float32)
    idx, d2, found = orc.ref_knn(cloud, q, 5)
    tiny = cloud[:3]
    idx_t, d2_t, found_t = orc.ref_knn(tiny, q[:10], 5)
    np.savez_compressed(os.path.join(OUT, "ref_knn.npz"), cloud=cloud, q=q, idx=idx, d2=d2, found=found, idx_t=idx_t, found_t=found_t)
    # --- real reference map + ICP (libref_core.so)
    from oracle import ref
    assert ref.available(), "oracle/_ref/libref_core.so missing"
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_reference_core import golden_run
    np.savez_compressed(os.path.join(OUT, "ref_core.npz"), **golden_run(orc, ref))
    # --- oracle known answers
    scans, poses = synth.kitti_sequence(n_scans=3, seed=3, n_rings=32, n_az=400)
    feat, keys = orc.voxel_filter(scans[0][:, :3], 8, 0.5)
    pipe = orc.Pipeline()
    res = [pipe.process(s) for s in scans]
    m = pipe.map()
    k0, c0, n0 = m.export_l0()
    np.savez_compressed(os.path.join(OUT, "oracle_kat.npz"), feat=feat, keys=keys, poses=np.stack([r["pose"] for r in res]),
                        n_corr=np.array([r["n_corr"] for r in res]), l0_keys=k0, l0_cent=c0, l0_cnt=n0,
                        shuffle_head_1000=orc.shuffle_head(1000, 100), shuffle_head_65536=orc.shuffle_head(65536, 100),
                        shuffle_head_4321=orc.shuffle_head(4321, 100))
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
