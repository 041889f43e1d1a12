"""Times the reference's OWN processing::Estimator::process_frame (the unmodified Estimator.cpp built against the drop-in shim and
libb2lo.so: oracle/_ref/libref_estimator_gpu.so) on a recorded sequence.  Run by bench.py as a child process (a crash here must not
take the bench line with it): argv = <scans.bin> <warmup>; the file holds, per scan, a u32 point count and the xyzi f32 records.
Prints one JSON line: wall ms per scan, host clock, pageable clouds, loop detection and pose graph switched off."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import orc, ref

path, W = sys.argv[1], int(sys.argv[2])
raw = np.fromfile(path, np.uint8)
scans, off = [], 0
while off < raw.size:
    n = int(raw[off:off + 4].view(np.uint32)[0]); off += 4
    scans.append(raw[off:off + 16 * n].view(np.float32).reshape(n, 4).copy()); off += 16 * n
est = ref.Estimator(gpu=True)
for s in scans[:W]:
    est.process(s)
kf = 0
inner = 0.0
t0 = time.perf_counter()
for s in scans[W:]:
    kf += int(est.process(s)["keyframe"])
    try:
        inner += est.last_process_ms() or 0.0
    except Exception:
        inner = float("nan")
dt = time.perf_counter() - t0
K = len(scans) - W
l0, l1, ns = est.counts()
print(json.dumps({"scans": K, "ms_per_scan": 1e3 * dt / K, "scans_per_s": K / dt, "process_frame_ms_per_scan": (inner / K) if inner == inner and inner > 0 else None,
                  "keyframes": kf, "map_l0": l0, "map_l1": l1,
                  "note": "ms_per_scan includes the binding's conversion of the flat scan into a util::PointCloud (the player's job in the reference); process_frame_ms_per_scan is Estimator::process_frame alone"}))
