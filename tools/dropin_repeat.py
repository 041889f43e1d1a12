"""Debug aid (GPU box): the drop-in bench (shim/test/dropin_bench) several times in a row on one box with the map-update trace
(B2LO_TRACE_UPDATE) on, to see where the run-to-run variance of its UpdateVoxelMap stage comes from.  One summary line per run: the
stage means of the program and, from the trace, the in-call times of the slowest staging copy / reserve / update."""
import json, os, re, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

scans, _ = bench.make_scans(106, 42, "cuda:0")
bench.dropin_leg(scans[:12], 5, 5)      # builds the executable
exe = os.path.join(ROOT, "lidar_odometry_b200", "shim", "test", "dropin_bench")
os.environ["B2LO_TRACE_UPDATE"] = "1"
for i in range(5):
    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
        for s in scans[:105]:
            a = np.ascontiguousarray(s[:, :4], np.float32)
            f.write(np.uint32(a.shape[0]).tobytes()); f.write(a.tobytes())
        path = f.name
    p = subprocess.run([exe, path, "5"], capture_output=True, text=True)
    os.unlink(path)
    out = json.loads(p.stdout.strip().splitlines()[-1])
    tr = [(float(m.group(1)), float(m.group(2)), float(m.group(3)))
          for m in re.finditer(r"stage ([0-9.]+) us, reserve ([0-9.]+) us \([^)]*\), update ([0-9.]+) us", p.stderr)]
    ups = sorted(t[2] for t in tr)
    print(i, "ms/scan", out["ms_per_scan"], "update stage", out["stage_ms_per_scan"]["update_voxel_map"], "| in-call update us: median",
          ups[len(ups) // 2] if ups else None, "max", ups[-1] if ups else None, "| stage max", max(t[0] for t in tr) if tr else None,
          "reserve max", max(t[1] for t in tr) if tr else None, flush=True)
