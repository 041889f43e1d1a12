"""Debug aid (GPU box): in-graph timeline of one LOCK-STEP step (S sequences, one batched kernel per step of the scan).  Needs the
instrumented build: B2LO_LIB=lidar_odometry_b200/libb2lo_tl.so python tools/gpu_timeline_lockstep.py 128"""
import collections, ctypes as C, os, re, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from lidar_odometry_b200 import api, capi, synth
S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
STEPS = 24
FILES = {1: "b2lo_filter.cu", 2: "b2lo_icp.cu", 3: "b2lo_odom.cu", 4: "b2lo_map.cu"}
src = {i: open(os.path.join(ROOT, "lidar_odometry_b200", "csrc", f)).read().split("\n") for i, f in FILES.items()}
def label(fid, line):
    text = src[fid][line - 1]
    if "TL_HERE" in text:
        return "  . " + text.split("//")[-1].strip()
    for k in range(line - 1, max(line - 12, -1), -1):
        m = re.search(r"\bstruct\s+(k_[a-z0-9_]+)", src[fid][k])
        if m:
            return m.group(1)
    return f"{FILES[fid]}:{line}"
L = capi.lib()
L.b2lo_debug_timeline.restype = C.c_int
n_have = STEPS + 8
scans, _ = synth.kitti_sequence(n_scans=n_have, seed=42, device="cuda")
dev = [torch.from_numpy(np.ascontiguousarray(s)).cuda() for s in scans]
odos = [api.Odometry(api.Context(0)) for _ in range(S)]
ls = api.LockstepBatch(odos)
CAP = 1 << 20
buf = (C.c_ulonglong * (2 * CAP))()
agg = collections.defaultdict(list)
tot = []
for k in range(STEPS + 6):
    torch.cuda.synchronize()
    L.b2lo_debug_timeline(buf, CAP)
    res, ms = ls.process_dev([(dev[(k + j) % n_have].data_ptr(), scans[(k + j) % n_have].shape[0]) for j in range(S)], 4)
    n = L.b2lo_debug_timeline(buf, CAP)
    if k < 6:
        continue
    marks = sorted(((buf[2 * i + 1], label(buf[2 * i] >> 32, buf[2 * i] & 0xffffffff)) for i in range(n)))
    marks = [m for m in marks if not m[1].startswith("  .")]
    # consecutive marks of one kernel = one batched launch: its first and last sequence start
    groups = []
    for t, name in marks:
        if groups and groups[-1][0] == name:
            groups[-1][2] = t
        else:
            groups.append([name, t, t])
    t0 = groups[0][1]
    for g, nxt in zip(groups, groups[1:] + [[None, groups[-1][2], 0]]):
        agg[g[0]].append((nxt[1] - g[1]) * 1e-3)
    tot.append(((groups[-1][2] - t0) * 1e-3, ms * 1e3))
    if k == 6:
        for g, nxt in zip(groups, groups[1:] + [[None, groups[-1][2], 0]]):
            print(f"{(g[1] - t0) * 1e-3:9.1f} us  +{(nxt[1] - g[1]) * 1e-3:8.1f}  (starts spread {(g[2] - g[1]) * 1e-3:7.1f})  {g[0]}")
print(f"\n=== S = {S}: mean interval per kernel over {len(tot)} steps")
for name, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{name:22s} launches/step {len(v) / len(tot):5.2f}  mean {np.mean(v):8.1f} us  per step {sum(v) / len(tot):8.1f} us")
print(f"first-to-last mark {np.mean([t[0] for t in tot]):.1f} us, CUDA-event time per step {np.mean([t[1] for t in tot]):.1f} us")
