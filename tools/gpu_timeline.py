"""Debug aid (GPU box): the in-graph timeline of one steady-state scan.  Needs the instrumented build (make -C lidar_odometry_b200/csrc
TIMELINE=1 -> libb2lo_tl.so), selected through B2LO_LIB.  Every kernel of the per-scan path stamps %globaltimer when its first thread
starts; the start-to-start intervals are what a kernel and the boundary behind it cost INSIDE the replayed CUDA graph (ncu cannot see
into a graph replay, and its serialised launch list says nothing about the gaps).

    B2LO_LIB=lidar_odometry_b200/libb2lo_tl.so python tools/gpu_timeline.py [--lookahead] [--scans 40]
"""
import argparse, collections, ctypes as C, os, re, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from lidar_odometry_b200 import api, capi, synth

ap = argparse.ArgumentParser()
ap.add_argument("--lookahead", action="store_true")
ap.add_argument("--scans", type=int, default=40)
ap.add_argument("--no-flush", action="store_true", help="leave L2 warm between scans (bounds what prefetching could buy)")
ap.add_argument("--dump", type=int, default=1, help="print the full timeline of this many scans")
args = ap.parse_args()

FILES = {1: "b2lo_filter.cu", 2: "b2lo_icp.cu", 3: "b2lo_odom.cu", 4: "b2lo_map.cu"}
src = {i: open(os.path.join(ROOT, "lidar_odometry_b200", "csrc", f)).read().split("\n") for i, f in FILES.items()}


def label(fid, line):
    """kernel name of a TL_START mark (searched upwards from the mark), or the trailing comment of a TL_HERE mark"""
    text = src[fid][line - 1]
    if "TL_HERE" in text:
        return "  . " + text.split("//")[-1].strip()
    for k in range(line - 1, max(line - 12, -1), -1):
        m = re.search(r"\bstruct\s+(k_[a-z0-9_]+)", src[fid][k]) or re.search(r"\b(k_[a-z0-9_]+)\s*\(", src[fid][k])
        if m:
            return m.group(1)
    return f"{FILES[fid]}:{line}"


L = capi.lib()
L.b2lo_debug_timeline.restype = C.c_int
scans, _ = synth.kitti_sequence(n_scans=args.scans + 6, seed=42, device="cuda")
dev = [torch.from_numpy(np.ascontiguousarray(s)).cuda() for s in scans]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
odo = api.Odometry()
buf = (C.c_ulonglong * (2 * 8192))()
agg = collections.defaultdict(list)
totals = []
for i in range(args.scans + 5):
    t, u = dev[i], dev[i + 1]
    if not args.no_flush:
        flush.zero_()
    torch.cuda.synchronize()
    L.b2lo_debug_timeline(buf, 8192)          # clear
    la = (u.data_ptr(), u.shape[0], u.stride(0)) if args.lookahead else None
    r = odo.process_dev(t.data_ptr(), t.shape[0], t.stride(0), lookahead=la)
    n = L.b2lo_debug_timeline(buf, 8192)
    if i < 5:
        continue
    marks = sorted(((buf[2 * k + 1], buf[2 * k] >> 32, buf[2 * k] & 0xffffffff) for k in range(n)))
    t0 = marks[0][0]
    rows = [(tm - t0, label(fid, ln)) for tm, fid, ln in marks]
    totals.append((rows[-1][0] * 1e-3, r["device_ms"] * 1e3, r["n_iters"]))
    for (a, name), (b, _) in zip(rows, rows[1:] + [(rows[-1][0], "")]):
        agg[name].append((b - a) * 1e-3)
    if i - 5 < args.dump:
        print(f"--- scan {i}: {r['n_iters']} GN iterations, {r['n_corr']} correspondences, event time {r['device_ms'] * 1e3:.1f} us")
        for (a, name), (b, _) in zip(rows, rows[1:] + [(rows[-1][0], "")]):
            print(f"{a * 1e-3:9.2f} us  +{(b - a) * 1e-3:7.2f}  {name}")
print("\n=== mean start-to-next-start interval per mark over", len(totals), "scans (us per scan = mean x count / scans)")
tot = 0.0
for name, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    per_scan = sum(v) / len(totals)
    tot += per_scan
    print(f"{name:28s} n/scan {len(v) / len(totals):5.2f}  mean {np.mean(v):7.2f} us  per scan {per_scan:7.2f} us")
k1 = sum(sum(v) for n, v in agg.items() if n.startswith("k_flt")) / len(totals)
print(f"K1 {k1:.1f} us per scan")
print(f"sum {tot:.1f} us; first-mark-to-last-mark {np.mean([t[0] for t in totals]):.1f} us; CUDA-event time {np.mean([t[1] for t in totals]):.1f} us; GN iterations {np.mean([t[2] for t in totals]):.2f}")
