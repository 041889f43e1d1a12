"""Scan ingest: the reference's loaders (util::load_kitti_binary, PLYPlayer::parse_ply_header / load_ply_point_cloud)
against the oracle restatement on file images, and K1 reading the file's own records in place against the oracle filter of the
loaded cloud - bit for bit."""
import numpy as np
import pytest

from lidar_odometry_b200 import api


def _ply(props, n, body=b"", fmt="binary_little_endian", pre="", extra=""):
    lines = ["ply", f"format {fmt} 1.0", "comment made by tests"]
    if pre:
        lines.append(pre)
    lines.append(f"element vertex {n}")
    lines += [f"property {t} {name}" for t, name in props]
    if extra:
        lines.append(extra)
    lines.append("end_header")
    return ("\n".join(lines) + "\n").encode() + body


def _records(xyz, layout):
    """layout: list of (kind, payload) building one record per point: 'x','y','z' floats, or ('pad', k) filler bytes."""
    rng = np.random.default_rng(3)
    n = xyz.shape[0]
    cols = []
    for kind in layout:
        if kind in ("x", "y", "z"):
            cols.append(np.ascontiguousarray(xyz[:, "xyz".index(kind)].astype(np.float32)).view(np.uint8).reshape(n, 4))
        else:
            cols.append(rng.integers(0, 256, size=(n, kind[1]), dtype=np.uint8))
    return np.ascontiguousarray(np.concatenate(cols, axis=1))


CASES = {
    "xyz": ([("float", "x"), ("float", "y"), ("float", "z")], ["x", "y", "z"]),
    "xyz_rgb_15B": ([("float", "x"), ("float", "y"), ("float", "z"), ("uchar", "red"), ("uchar", "green"), ("uchar", "blue")],
                    ["x", "y", "z", ("pad", 3)]),
    "time_first": ([("double", "t"), ("float32", "x"), ("float32", "y"), ("float32", "z"), ("ushort", "ring")],
                   [("pad", 8), "x", "y", "z", ("pad", 2)]),
    "zyx_order": ([("uchar", "flag"), ("float", "z"), ("float", "y"), ("short", "s"), ("float", "x")],
                  [("pad", 1), "z", "y", ("pad", 2), "x"]),
    "unknown_type_counts_4": ([("float", "x"), ("weird", "w"), ("float", "y"), ("float", "z")], ["x", ("pad", 4), "y", "z"]),
}


def _cloud(n=4000, seed=1):
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-40, 40, size=(n, 3)).astype(np.float32)
    xyz[:, 2] = rng.uniform(-2, 6, size=n).astype(np.float32)
    xyz[5] = [np.nan, 1, 2]      # non-finite points are skipped by the filter, not by the loader
    xyz[min(7, n - 1)] = [3, np.inf, 2]
    return xyz


@pytest.mark.parametrize("case", sorted(CASES))
def test_ply_header_and_binary_body_match_the_oracle(orc, case):
    props, layout = CASES[case]
    xyz = _cloud(257)
    rec = _records(xyz, layout)
    img = _ply(props, xyz.shape[0], rec.tobytes())
    h = api.parse_ply_header(img)
    assert h is not None and h["is_binary"] and h["vertex_count"] == 257 and h["n_records"] == 257
    assert h["fmt"].record_bytes == rec.shape[1]
    got = api.load_ply_point_cloud(img)
    ref = orc.ply_load(img)
    assert got.shape == ref.shape == (257, 3)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert np.array_equal(got.view(np.uint32), xyz.view(np.uint32))


def test_ply_header_quirks_follow_the_reference(orc):
    xyz = _cloud(10)
    rec = _records(xyz, ["x", "y", "z"])
    # properties of LATER elements are appended to the vertex record as well (ply_player.cpp:425-441): list -> 4 bytes
    img = _ply(CASES["xyz"][0], 10, rec.tobytes(), extra="element face 2\nproperty list uchar int vertex_indices")
    h = api.parse_ply_header(img)
    assert h["fmt"].record_bytes == 16 and h["n_records"] == 7          # 120 body bytes / 16
    assert np.array_equal(api.load_ply_point_cloud(img).view(np.uint32), orc.ply_load(img).view(np.uint32))
    # a truncated last record is dropped (:321-324)
    img = _ply(CASES["xyz"][0], 10, rec.tobytes()[:-5])
    assert api.parse_ply_header(img)["n_records"] == 9
    assert api.load_ply_point_cloud(img).shape == orc.ply_load(img).shape == (9, 3)
    # more vertices announced than present
    img = _ply(CASES["xyz"][0], 50, rec.tobytes())
    assert api.load_ply_point_cloud(img).shape == orc.ply_load(img).shape == (10, 3)
    # rejected files: no z, zero vertices, CRLF header (exact-match "ply" line), missing magic, bad count
    for bad in (_ply([("float", "x"), ("float", "y")], 10, rec.tobytes()), _ply(CASES["xyz"][0], 0), _ply(CASES["xyz"][0], 10).replace(b"\n", b"\r\n"),
                _ply(CASES["xyz"][0], 10, rec.tobytes())[4:], _ply(CASES["xyz"][0], "many", rec.tobytes()), b""):
        assert api.parse_ply_header(bad) is None
        assert api.load_ply_point_cloud(bad).shape == orc.ply_load(bad).shape == (0, 3)
    # no end_header: header accepted, nothing left to read
    img = _ply(CASES["xyz"][0], 10).replace(b"end_header\n", b"") + rec.tobytes()
    assert api.load_ply_point_cloud(img).shape == orc.ply_load(img).shape == (0, 3)
    # big-endian bodies are copied without a byte swap, as the reference does
    img = _ply(CASES["xyz"][0], 10, rec.tobytes(), fmt="binary_big_endian")
    assert np.array_equal(api.load_ply_point_cloud(img).view(np.uint32), xyz.view(np.uint32))
    # duplicate coordinate names: the last one wins
    img = _ply([("float", "x"), ("float", "x"), ("float", "y"), ("float", "z")], 10, _records(xyz, [("pad", 4), "x", "y", "z"]).tobytes())
    assert api.parse_ply_header(img)["fmt"].off_x == 4
    assert np.array_equal(api.load_ply_point_cloud(img).view(np.uint32), orc.ply_load(img).view(np.uint32))


def test_ply_ascii_body_matches_the_oracle(orc):
    body = "\n".join([
        "1 2 3 255", "4.5 -6.25 7e-2 0", "  8\t9   10  1  ", "11 12", "", "13 14 15 16 17 18", "1e5 .5 -.25 3", "19 2x 21 22", "+1 -2 +3.5e+1 0",
        "nan 1 2 3", "23 24 25 26", "0x10 1 2 3", "1e 2 3 4", "27 28 29 30",
    ]) + "\n"
    img = _ply([("float", "x"), ("float", "y"), ("float", "z"), ("uchar", "i")], 14, body.encode(), fmt="ascii")
    h = api.parse_ply_header(img)
    assert h is not None and not h["is_binary"]
    got, ref = api.load_ply_point_cloud(img), orc.ply_load(img)
    assert got.shape == ref.shape and got.shape[0] >= 8
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert got[0].tolist() == [1.0, 2.0, 3.0] and got[1].tolist() == [4.5, -6.25, np.float32(7e-2)]
    # fewer lines than vertices: stops at the end of the file
    img = _ply(CASES["xyz"][0], 100, b"1 2 3\n4 5 6\n", fmt="ascii")
    assert api.load_ply_point_cloud(img).shape == orc.ply_load(img).shape == (2, 3)


def test_kitti_image_matches_the_oracle(orc, tmp_path):
    xyz = _cloud(1000)
    rec = np.concatenate([xyz, np.ones((1000, 1), np.float32)], axis=1)
    p = tmp_path / "000000.bin"
    p.write_bytes(rec.tobytes() + b"\x01\x02\x03\x04\x05")      # a trailing partial record is ignored (PointCloudUtils.cpp:42)
    got = api.load_kitti_binary(str(p))
    ref = orc.kitti_load(p.read_bytes())
    assert got.shape == (1000, 4)
    assert np.array_equal(np.ascontiguousarray(got[:, :3]).view(np.uint32), ref.view(np.uint32))
    f = api.kitti_record_format()
    assert (f.record_bytes, f.off_x, f.off_y, f.off_z) == (16, 0, 4, 8)


# ---- GPU: K1 over the file's own records --------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("pinned", [False, True])
def test_filter_reads_ply_records_in_place(b2, orc, case, pinned):
    import torch
    props, layout = CASES[case]
    xyz = _cloud(20000, seed=4)
    rec = _records(xyz, layout)
    img = np.frombuffer(_ply(props, xyz.shape[0], rec.tobytes()), dtype=np.uint8)
    h = b2.parse_ply_header(img)
    if pinned:
        img = torch.from_numpy(img.copy()).pin_memory().numpy()     # page-locked image: K1 reads the sampled records over PCIe in place
    f = b2.FastVoxelFilter(0.5)
    for stride in (1, 4):
        got = f.filter_records(img, h["fmt"], n_records=h["n_records"], stride=stride, want_keys=True, offset=h["data_offset"])
        ref, keys = orc.voxel_filter(orc.ply_load(img), stride, 0.5)
        assert got.shape == ref.shape and f.getVoxelCount() == ref.shape[0]
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
        assert np.array_equal(f.last_keys, keys)
    assert f.filter_records(img[: h["data_offset"]], h["fmt"], n_records=0, offset=h["data_offset"]).shape == (0, 3)


@pytest.mark.gpu
def test_odometry_on_kitti_file_images_equals_odometry_on_clouds(b2, orc, small_kitti):
    """The per-scan driver fed with .bin file images (pageable, page-locked + look-ahead, device-resident) gives the poses and map of the
    driver fed with the loaded clouds, bit for bit; and PLY images of the same scans (19-byte records) do too."""
    import torch
    scans, _ = small_kitti
    images = [np.ascontiguousarray(s.astype(np.float32)).view(np.uint8).reshape(-1).copy() for s in scans]     # xyzI records = the .bin file
    base = b2.Odometry()
    want = [base.process(s) for s in scans]
    l0_want = base.map().export_l0()

    def same(res):
        for k, (a, b) in enumerate(zip(res, want)):
            assert np.array_equal(a["pose"].view(np.uint32), b["pose"].view(np.uint32)), k
            assert (a["keyframe"], a["n_features"], a["n_corr"], a["n_iters"]) == (b["keyframe"], b["n_features"], b["n_corr"], b["n_iters"]), k

    kf = b2.kitti_record_format()
    # pageable images
    odo = b2.Odometry()
    odo.set_record_format(kf)
    same([odo.process_records(im) for im in images])
    got = odo.map().export_l0()
    assert all(np.array_equal(np.asarray(x).view(np.uint8), np.asarray(y).view(np.uint8)) for x, y in zip(got, l0_want))
    # page-locked images with look-ahead
    pins = [torch.from_numpy(im).pin_memory() for im in images]
    odo = b2.Odometry()
    odo.set_record_format(kf)
    res = []
    for k, pin in enumerate(pins):
        la = (pins[k + 1].numpy(), pins[k + 1].numel() // 16) if k + 1 < len(pins) else None
        res.append(odo.process_records(pin.numpy(), lookahead=la))
    same(res)
    # device-resident images
    devs = [torch.from_numpy(im).cuda() for im in images]
    odo = b2.Odometry()
    odo.set_record_format(kf)
    same([odo.process_dev(d.data_ptr(), d.numel() // 16, 3) for d in devs])
    # the same scans as PLY files with unaligned 19-byte vertices
    props = [("uchar", "flag"), ("float", "x"), ("float", "y"), ("float", "z"), ("ushort", "ring"), ("float", "intensity")]
    odo = b2.Odometry()
    res = []
    for s in scans:
        img = np.frombuffer(_ply(props, s.shape[0], _records(s[:, :3], [("pad", 1), "x", "y", "z", ("pad", 6)]).tobytes()), dtype=np.uint8)
        h = b2.parse_ply_header(img)
        odo.set_record_format(h["fmt"])
        pin = torch.from_numpy(img.copy()).pin_memory().numpy()
        res.append(odo.process_records(pin, n_records=h["n_records"], offset=h["data_offset"]))
    same(res)
    # back to float clouds
    odo.set_record_format(None)
    assert odo.process(scans[-1])["ok"]


# ---- final-map export: util::VoxelGrid ------------------------------------------------------------------
def test_oracle_voxel_grid_known_answers(orc):
    x = np.array([[0.1, 0.1, 0.1], [0.2, 0.2, 0.2], [-0.1, 5, 5], [0.3, 0.3, 0.3], [0.9, -0.2, 0.0]], np.float32)
    got = orc.voxel_grid_filter(x, 1.0)
    # std::map order: (-1,5,5) < (0,-1,0) < (0,0,0); the (0,0,0) cell holds the running mean of 3 points
    c = np.float32(0.5) * x[0] + np.float32(0.5) * x[1]
    c = (np.float32(2.0) / np.float32(3.0)) * c + (np.float32(1.0) / np.float32(3.0)) * x[3]
    assert got.shape == (3, 3)
    assert np.array_equal(got[0], x[2]) and np.array_equal(got[1], x[4])
    assert np.array_equal(got[2].view(np.uint32), c.astype(np.float32).view(np.uint32))
    assert orc.voxel_grid_filter(x, 0.0).shape == (0, 3) and orc.voxel_grid_filter(x[:0], 1.0).shape == (0, 3)


@pytest.mark.gpu
def test_voxel_grid_export_matches_the_oracle(b2, orc, small_kitti):
    scans, poses = small_kitti
    # what save_map_to_ply accumulates: keyframe feature clouds moved to world coordinates (Estimator.cpp:1262-1276)
    f = b2.FastVoxelFilter(0.5)
    world = []
    for s, T in zip(scans, poses):
        feat = f.filter(s, 8)
        T = np.asarray(T, np.float32)
        world.append((feat @ T[:3, :3].T + T[:3, 3]).astype(np.float32))
    acc = np.ascontiguousarray(np.concatenate(world))
    rng = np.random.default_rng(8)
    dense = rng.normal(0, 0.6, size=(30000, 3)).astype(np.float32)              # many points per voxel (> 32), negative coordinates
    for cloud, leaf in ((acc, 0.4), (acc, 1.0), (dense, 0.25), (dense[:1], 0.1)):
        g = b2.VoxelGrid()
        g.setLeafSize(leaf)
        g.setInputCloud(cloud)
        got = g.filter()
        ref = orc.voxel_grid_filter(cloud, leaf)
        assert got.shape == ref.shape and got.shape[0] > 0
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    g = b2.VoxelGrid()
    g.setLeafSize(0.0)
    g.setInputCloud(acc)
    assert g.filter().shape == (0, 3)
    # the feature path still works after the export reused the K1 buffers
    again = f.filter(scans[0], 8)
    ref0, _ = orc.voxel_filter(scans[0][:, :3], 8, 0.5)
    assert np.array_equal(again.view(np.uint32), ref0.view(np.uint32))
