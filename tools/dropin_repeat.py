"""Debug aid (GPU box): the drop-in bench several times in a row on one box, with the map-update trace (B2LO_TRACE_UPDATE) on, to
see where the run-to-run variance of its UpdateVoxelMap stage comes from.  Output: one summary line per run + the slowest traced updates."""
import os, sys, json, subprocess, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
scans, _ = bench.make_scans(106, 42, "cuda:0")
os.environ["B2LO_TRACE_UPDATE"] = "1"
import io, contextlib
for i in range(5):
    # run the compiled bench directly so that stderr can be captured per run
    r = bench.dropin_leg(scans, 100, 5, show_stderr=False) if False else None
    import tempfile, numpy as np
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "lidar_odometry_b200", "shim", "test", "dropin_bench")
    if i == 0:
        bench.dropin_leg(scans[:12], 5, 5)      # builds the executable
    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
        for s in scans[:105]:
            a = np.ascontiguousarray(s[:, :4], np.float32)
            f.write(np.uint32(a.shape[0]).tobytes()); f.write(a.tobytes())
        path = f.name
    p = subprocess.run([exe, path, "5"], capture_output=True, text=True)
    os.unlink(path)
    out = json.loads(p.stdout.strip().splitlines()[-1])
    tr = [(float(m.group(1)), float(m.group(2)), float(m.group(3))) for m in re.finditer(r"stage ([0-9.]+) us, reserve ([0-9.]+) us \([^)]*\), update ([0-9.]+) us", p.stderr)]
    ups = sorted(t[2] for t in tr)
    print(i, "ms/scan", out["ms_per_scan"], "update stage", out["stage_ms_per_scan"]["update_voxel_map"], "| in-call update us: median", ups[len(ups) // 2] if ups else None,
          "max", ups[-1] if ups else None, "sum ms", round(sum(ups) / 1e3, 2), "| stage max", max(t[0] for t in tr) if tr else None, "reserve max", max(t[1] for t in tr) if tr else None, flush=True)
