"""Debug aid (GPU box): lock-step throughput of single batches (CUDA events per step) and of batches driven from several host threads
(wall clock, host work included)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from lidar_odometry_b200 import api
K, W = 40, 5
scans, _ = bench.make_scans(K + W + 1, 42, "cuda:0")
dev = [torch.from_numpy(s).cuda() for s in scans]
def dev_args(i):
    return dev[i].data_ptr(), scans[i].shape[0], scans[i].shape[1]
cfg = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(1, 384), (3, 128), (2, 192), (4, 96)]
for G, S in cfg:
    r = bench.lockstep_groups_leg(api, 0, dev_args, G, S, K, W)
    print(f"groups {G}x{S}:", round(r["scans_per_s"]), "scans/s (wall clock)", flush=True)
