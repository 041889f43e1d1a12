import json, os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch, bench
from lidar_odometry_b200 import api
K, W = 40, 5
scans, _ = bench.make_scans(K + W + 1, 42, "cuda:0")
dev = [torch.from_numpy(s).cuda() for s in scans]
def dev_args(i):
    return dev[i].data_ptr(), scans[i].shape[0], scans[i].shape[1]
for S in (128, 256):
    r = bench.lockstep_leg(api, 0, dev_args, S, K, W, 0, "")
    print(f"lockstep {S}:", round(r["scans_per_s"]), "scans/s", flush=True)
for G, S in [(3, 96), (4, 96), (2, 128)]:
    r = bench.lockstep_groups_leg(api, 0, dev_args, G, S, K, W)
    print(f"groups {G}x{S}:", round(r["scans_per_s"]), "scans/s", flush=True)
