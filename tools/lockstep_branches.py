"""Debug aid (GPU box): lock-step throughput of one batch against the number of branches of its graph (B2LO_LOCKSTEP_BRANCHES)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from lidar_odometry_b200 import api
K, W = 40, 5
scans, _ = bench.make_scans(K + W + 1, 42, "cuda:0")
dev = [torch.from_numpy(s).cuda() for s in scans]
def dev_args(i):
    return dev[i].data_ptr(), scans[i].shape[0], scans[i].shape[1]
for S, Bs in ((128, (2, 4, 8)), (256, (4, 6, 8)), (384, (6, 8))):
    for B in Bs:
        os.environ["B2LO_LOCKSTEP_BRANCHES"] = str(B)
        r = bench.lockstep_leg(api, 0, dev_args, S, K, W, 0, "")
        print(f"lockstep {S} sequences, {B} branches:", round(r["scans_per_s"]), "scans/s", flush=True)
