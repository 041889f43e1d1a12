"""GPU box, under ncu (--graph-profiling node): a few lock-step steps at S sequences, for the batched K2 kernel's counters."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lidar_odometry_b200 import api, synth
S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
scans, _ = synth.kitti_sequence(n_scans=12, seed=42, device="cuda")
dev = [torch.from_numpy(np.ascontiguousarray(s)).cuda() for s in scans]
odos = [api.Odometry(api.Context(0)) for _ in range(S)]
ls = api.LockstepBatch(odos)
nq = 0
for k in range(8):
    res, ms = ls.process_dev([(dev[(k + j) % 12].data_ptr(), scans[(k + j) % 12].shape[0]) for j in range(S)], 4)
    nq = sum(r["n_features"] for r in res)
print("ok", S, "sequences,", nq, "queries in the last step's first iteration,", ms, "ms")
